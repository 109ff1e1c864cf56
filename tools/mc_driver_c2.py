"""Full-size run of the product Monte-Carlo driver: writes a C2-like world (512^3 BOSS-like n(r), small catalogues that only
fix alpha, the shot factor and the box centre) under a scratch directory and times
`python -m spectra_without_windows_b200.mc_driver pk <param>` for N_mc realisations, device RNG included."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spectra_without_windows_b200 import synthetic as syn  # noqa: E402

n_mc = int(sys.argv[1]) if len(sys.argv) > 1 else 12
extra = sys.argv[2:]
nproc = 1
if "--nproc" in extra:          # torchrun over several GPUs of the node
    i = extra.index("--nproc")
    nproc = int(extra[i + 1])
    extra = extra[:i] + extra[i + 2:]
wl = bench.WORKLOADS["C2"]
grid, box, bc = wl["grid"], wl["box"], wl["box_center"]
rax = syn.mesh_axes(box, grid, bc)
nbar_r = bench.survey_nbar_r([np.asarray(a) for a in rax], np)
rng = np.random.default_rng(3)
lo, hi = np.asarray(bc) - 0.5 * np.asarray(box), np.asarray(bc) + 0.5 * np.asarray(box)
# the catalogues only enter through alpha, the shot-noise factor and the box centre of the randoms
rpos = np.vstack([rng.uniform(lo, hi, size=(49998, 3)), lo[None, :], hi[None, :]])
dpos = rng.uniform(lo, hi, size=(1000, 3))
w = syn.World(name="c2", grid=grid, box=box, box_center=bc, type="c2_", pk_bins=wl["bins"], N_mc=n_mc)
w.data = {"pos": dpos, "w": np.ones(len(dpos))}
w.randoms = {"pos": rpos, "w": np.ones(len(rpos))}
w.nbar_r = nbar_r
root = tempfile.mkdtemp(prefix="sww_c2_")
p = syn.write_world_files(w, root)
t0 = time.time()
launch = [sys.executable, "-m", "spectra_without_windows_b200.mc_driver"]
if nproc > 1:
    launch = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
              "--master-port", "29541", "-m", "spectra_without_windows_b200.mc_driver"]
r = subprocess.run(launch + ["pk", p] + extra, capture_output=True, text=True, cwd=ROOT)
dt = time.time() - t0
print(r.stdout[-600:], r.stderr[-1500:])
assert r.returncode == 0
F = np.load([os.path.join(root, "out", f) for f in os.listdir(os.path.join(root, "out")) if f.startswith("fisher-pk_")][0])
print("mc_driver pk C2 on %d GPU(s): N_mc=%d %s wall %.1f s (includes start-up: context, n(r) load, Y_lm cache) -> F %s finite=%s"
      % (nproc, n_mc, " ".join(extra), dt, F.shape, bool(np.isfinite(F).all())))
