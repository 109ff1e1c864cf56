"""In-tree build of libsww.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m spectra_without_windows_b200.build [--force]

Objects land in csrc/build/, the shared library next to this file (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsww.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]
SOURCES = ["ctx.cu", "kernels.cu", "fft.cu", "fft_sm100.cu", "pipelines.cu", "bk.cu", "gram.cu", "rng.cu", "ml.cu", "rowmesh.cu", "paint.cu"]


def _stamp(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "sww.h"))
    stamp = _stamp(deps)
    sfile = os.path.join(bdir, "stamp")
    if not force and os.path.exists(OUT) and os.path.exists(sfile) and open(sfile).read() == stamp:
        return OUT
    logs = {}

    def cc(src):
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, logs[src]))
        return obj

    with ThreadPoolExecutor(max_workers=min(12, len(SOURCES))) as ex:
        objs = list(ex.map(cc, SOURCES))
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        for s in SOURCES:
            f.write("==== %s\n%s\n" % (s, logs[s]))
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-L/usr/local/cuda/lib64", "-lcufft",
                                                  "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(sfile, "w").write(stamp)
    if verbose:
        print(open(os.path.join(bdir, "ptxas.log")).read())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
