"""Painter benchmark: tile-binned deterministic painter (csrc/paint.cu) against the first version (one thread per particle,
global FP64 atomics) on 1.5e7 clustered randoms drawn from the C2 survey n(r), TSC and CIC, at 128^3 and 512^3.
Prints one JSON line per case (ms per paint, particles/s).     python tools/paint_bench.py [n_particles]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import survey_nbar_r
    from spectra_without_windows_b200 import engine
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 15_000_000
    dev = torch.device("cuda", 0)
    box, bc = (1800.0, 3450.0, 1900.0), (1200.0, 0.0, 0.0)
    # clustered randoms: rejection-sample the survey selection function on a fine lattice of candidate points
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    chunks, have = [], 0
    boxt, bct = torch.tensor(box, device=dev, dtype=torch.float64), torch.tensor(bc, device=dev, dtype=torch.float64)
    while have < n:
        cand = (torch.rand((4_000_000, 3), generator=gen, device=dev, dtype=torch.float64) - 0.5) * boxt + bct
        X, Y, Z = cand[:, 0], cand[:, 1], cand[:, 2]
        r = torch.sqrt(X * X + Y * Y + Z * Z)
        lon, lat = torch.atan2(Y, X), torch.asin(Z / r)
        W = (lon.abs() < np.deg2rad(70.0)) * (lat.abs() < np.deg2rad(30.0)) * (0.9 + 0.1 * torch.cos(3.0 * lon))
        sel = torch.exp(-0.5 * ((r - 1400.0) / 250.0) ** 2) * (r > 1150.0) * (r < 1750.0) * W
        keep = torch.rand(cand.shape[0], generator=gen, device=dev, dtype=torch.float64) < sel
        chunks.append(cand[keep])
        have += int(keep.sum())
    pos = torch.cat(chunks)[:n].contiguous()
    w = torch.rand(n, generator=gen, device=dev, dtype=torch.float64) + 0.5
    for g in (128, 512):
        mesh = engine.Mesh(box, (g, g, g), bc, 0.0, 0.05, 0.025, 0)
        ctx = engine.Context(mesh, 0, ("ops",))
        for scheme in ("TSC", "CIC"):
            res = {}
            for tag, env in (("tile_binned", None), ("v1_global_atomics", "1")):
                if env:
                    os.environ["SWW_PAINT_V1"] = env
                try:
                    out = ctx.zeros(mesh.shape)
                    for _ in range(2):
                        ctx.paint(pos, w, bc, scheme=scheme, out=out)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    reps = 5
                    for _ in range(reps):
                        ctx.paint(pos, w, bc, scheme=scheme, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    a = ctx.paint(pos, w, bc, scheme=scheme)
                    b = ctx.paint(pos, w, bc, scheme=scheme)
                    res[tag] = {"ms": ms, "particles_per_s": n / (ms * 1e-3), "bit_identical_runs": bool(torch.equal(a, b)),
                                "sum_rel_err": float(abs(a.sum().item() - w.sum().item()) / w.sum().item())}
                    res[tag + "_map"] = a
                finally:
                    os.environ.pop("SWW_PAINT_V1", None)
            d = (res["tile_binned_map"] - res["v1_global_atomics_map"]).abs().max().item() / res["v1_global_atomics_map"].abs().max().item()
            print(json.dumps({"bench": "paint", "mesh": g, "scheme": scheme, "n_particles": n, "clustered": "C2 survey n(r)",
                              "tile_binned": res["tile_binned"], "v1_global_atomics": res["v1_global_atomics"],
                              "max_rel_diff_between_versions": d}))
        ctx.close()


if __name__ == "__main__":
    main()
