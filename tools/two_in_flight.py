"""Experiment: K independent P_l Fisher realisations in flight on one GPU (K contexts, K streams).  Prints realisations/s.
Measured on B200 at C2: K = 1 4.20/s, K = 2 4.42/s (the tails of one realisation's kernels overlap the other's).  Halving
every pass grid so that an FP64-bound z pass of one realisation shares each SM with an HBM-bound column pass of the
other did NOT help (4.16/s at K = 2, 3.95/s at K = 3 with a third of the slots each): the passes are latency bound and
occupancy is what they lack, so the idea was dropped."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spectra_without_windows_b200 import engine  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.WORKLOADS["C2"]
dev = torch.device("cuda", 0)
mesh = engine.Mesh(wl["box"], wl["grid"], wl["box_center"], *wl["bins"])
ctxs, streams, fields, outs = [], [], [], []
rax_d = [torch.from_numpy(a).to(dev) for a in mesh.rax]
nbar_r = bench.survey_nbar_r(rax_d, torch).contiguous()
nbar = (nbar_r * bench.ALPHA).contiguous()
nbar_A = nbar_r.clone()
nbar_A[nbar_A == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
sig = torch.sqrt(nbar_A)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
for k in range(K):
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        c = engine.Context(mesh, 0, ("pk_fisher",))
        c.set_fields(nbar, nbar, nbar_A, bench.SHOT)
    ctxs.append(c)
    streams.append(st)
    fields.append(torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64).mul_(sig))
    outs.append((torch.empty((c.n_b, c.n_b), dtype=torch.float64, device=dev), torch.empty((c.n_b,), dtype=torch.float64, device=dev)))
torch.cuda.synchronize()


def run(n):
    for i in range(n):
        for k in range(K):
            ctxs[k].pk_fisher(fields[k], outs[k][0], outs[k][1])
    torch.cuda.synchronize()


run(1)
t0 = time.perf_counter()
run(steps)
dt = time.perf_counter() - t0
print("K=%d: %.3f realisations/s (%.1f ms per realisation)" % (K, K * steps / dt, dt / (K * steps) * 1e3))
F = [o[0].clone() for o in outs]
assert all(bool(torch.isfinite(f).all()) for f in F)
