"""One 1-field-term accumulation (three trilinear Grams) at 256^3 -- target for `ncu -k regex:gram3`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectra_without_windows_b200 import engine

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mesh = engine.Mesh((1750.0, 3220.0, 1830.0), (256, 256, 256), (1200.0, 0.0, 0.0), 0.0, 0.11, 0.01, lmax)
ctx = engine.Context(mesh, 0, ("bk_fisher", "bk_data"))
n_l, n_k = mesh.n_l, mesh.n_k
g = torch.randn((n_l, n_k) + mesh.shape, dtype=torch.float64, device="cuda")
gm = torch.randn_like(g)
tg = torch.randn_like(g)
q = torch.zeros((ctx.n_tri, n_l), dtype=torch.float64, device="cuda")
for _ in range(2):
    ctx.bk_q1_accumulate(g, gm, tg, 10, q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ctx.bk_q1_accumulate(g, gm, tg, 10, q)
e1.record()
torch.cuda.synchronize()
print("q1_accumulate ms", e0.elapsed_time(e1), "finite", bool(torch.isfinite(q).all()))
