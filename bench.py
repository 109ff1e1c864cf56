"""Benchmark of the P_l(k) Fisher Monte-Carlo hot path (BASELINE.json metric:
"pk/bk Fisher MC realisations/s"), configuration C2 of SURVEY.md section 8d:
compute_pk_randoms.py Fisher matrix, BOSS-like synthetic n(r), 512^3 mesh, l<=4, dk=0.005.

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    (CPU: the oracle port on host cores)

A "step" is one Monte-Carlo realisation (one Gaussian field -> q-bar_alpha and F_ab).  Prints ONE
JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: grid, box, box_center, (k_min, k_max, dk, lmax)
    "C2": dict(grid=(512, 512, 512), box=(1800.0, 3450.0, 1900.0), box_center=(1200.0, 0.0, 0.0),
               bins=(0.0, 0.41, 0.005, 4)),
    "C5": dict(grid=(1024, 1024, 1024), box=(1800.0, 3450.0, 1900.0), box_center=(1200.0, 0.0, 0.0),
               bins=(0.0, 0.40, 0.005, 4), n_mc_job=200),
    "C1F": dict(grid=(128, 128, 128), box=(1000.0, 1000.0, 1000.0), box_center=(0.0, 0.0, 2000.0),
                bins=(0.0, 0.25, 0.01, 4)),
    # bispectrum Fisher (a step = one PAIR of realisations): 256^3, k in [0,0.11] dk=0.01, l<=4 -> 191 triangles, n_b=573
    "C4": dict(grid=(256, 256, 256), box=(1750.0, 3220.0, 1830.0), box_center=(1200.0, 0.0, 0.0),
               bins=(0.0, 0.11, 0.01, 4), bk=True),
    # bispectrum data path (config 3 of BASELINE.json): 256^3, k in [0,0.11] dk=0.01, l = 0,2; a step = one data catalogue's
    # g-maps + 3-field term + the 1-field term over N_mc = 10 resident bias simulations
    "C3": dict(grid=(256, 256, 256), box=(1750.0, 3220.0, 1830.0), box_center=(1200.0, 0.0, 0.0),
               bins=(0.0, 0.11, 0.01, 2), bk=True, bk_data=True, n_mc=100),
    # the Gram's comparison point: cuBLAS DGEMM on [573 x K] . [K x 573] (K = 2^22 cells of the 256^3 maps) against gram_dmma_kernel
    "DGEMM": dict(grid=(256, 256, 256), box=(1750.0, 3220.0, 1830.0), box_center=(1200.0, 0.0, 0.0),
                  bins=(0.0, 0.11, 0.01, 4), bk=True, dgemm=True),
    "C4S": dict(grid=(128, 128, 128), box=(1750.0, 3220.0, 1830.0), box_center=(1200.0, 0.0, 0.0),
                bins=(0.0, 0.11, 0.01, 2), bk=True),
    # pk data path (paint + q_alpha): config C1 of SURVEY 8d; a step = one catalogue
    "C1": dict(grid=(128, 128, 128), box=(1000.0, 1000.0, 1000.0), box_center=(0.0, 0.0, 2000.0),
               bins=(0.0, 0.25, 0.01, 4), data=True, n_gal=300000, n_ran=15000000),
    # the reference's production meshes (paramfiles/boss_nC.param:17-33): not powers of two -> any-length passes (fft_any.cuh)
    "BOSSPK": dict(grid=(258, 496, 274), box=(1800.0, 3450.0, 1900.0), box_center=(1200.0, 0.0, 0.0),
                   bins=(0.0, 0.41, 0.01, 4)),
    "BOSSBK": dict(grid=(172, 330, 182), box=(1800.0, 3450.0, 1900.0), box_center=(1200.0, 0.0, 0.0),
                   bins=(0.0, 0.11, 0.01, 4), bk=True),
    "S256": dict(grid=(256, 256, 256), box=(1800.0, 3450.0, 1900.0), box_center=(1200.0, 0.0, 0.0),
                 bins=(0.0, 0.20, 0.005, 4)),
}
ALPHA, SHOT = 0.02, 1.02


def survey_nbar_r(rax, xp):
    """Randoms-density map of SURVEY.md section 8d (C2): n(r) W_ang(r^) / alpha.  `xp` is numpy or torch."""
    X, Y, Z = xp.meshgrid(*rax, indexing="ij")
    r = xp.sqrt(X * X + Y * Y + Z * Z)
    lon = xp.arctan2(Y, X)
    lat = xp.arcsin(Z / (r + 1e-300))
    d70, d30 = np.deg2rad(70.0), np.deg2rad(30.0)
    W = (abs(lon) < d70) * (abs(lat) < d30) * (0.9 + 0.1 * xp.cos(3.0 * lon))
    n = 4e-4 * xp.exp(-0.5 * ((r - 1400.0) / 250.0) ** 2) * (r > 1150.0) * (r < 1750.0) * W
    return n / ALPHA


def uniform_nbar_r(rax, xp):
    X, _, _ = xp.meshgrid(*rax, indexing="ij")
    return X * 0.0 + 3e-4 / ALPHA


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def n_fft_min(M, n_b):
    return 2 * M + 1 + 2 * n_b


def oracle_axes(wl):
    """Cell-centre coordinates per axis with the oracle's restatement of load_coord_grids (src/opt_utilities.py:428-430);
    the reference arm must not lean on the product package."""
    from oracle import sww_oracle as orc
    box, grid = np.asarray(wl["box"], float), np.asarray(wl["grid"], int)
    off = np.asarray(wl["box_center"], float) + 0.5 * box / grid
    return [(orc._signed_index(grid[i]) * (box[i] / grid[i])).astype("f8") + off[i] for i in range(3)]


def reference_arm(wl, workload):
    from oracle.ref_arm import PkFisherRefArm
    nb_fn = uniform_nbar_r if workload == "C1F" else survey_nbar_r
    nbar_r = nb_fn(oracle_axes(wl), np)
    return PkFisherRefArm(wl["box"], wl["grid"], wl["box_center"], nbar_r, ALPHA, SHOT, *wl["bins"])


def run_reference(args):
    """CPU arm: the reference's own low_mem algorithm (pk/compute_pk_randoms.py:200-281) on the host cores, shipped
    precision, all threads; every step is a bounded sample of one realisation at the full mesh size (oracle/ref_arm.py).
    value = step_fraction / step time, where step_fraction is the share of one realisation's work a step covers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if wl.get("bk") or wl.get("data"):
        raise Exception("--impl reference times the pk Fisher workloads (C2, C5, C1F, S256)")
    arm = reference_arm(wl, args.workload)
    arm.prologue()
    for _ in range(args.warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    wall = time.perf_counter() - t0
    T = arm.T_real()          # means over all sampled rows (a kind of row not sampled yet is run once, untimed)
    parts = arm.parts()
    arm.close()
    ms_step = wall / args.steps * 1e3
    val = 1.0 / T
    grid = np.asarray(wl["grid"], int)
    n_k = int((wl["bins"][1] - wl["bins"][0]) / wl["bins"][2])
    n_b = n_k * (wl["bins"][3] // 2 + 1)
    cfg = {"workload": "%s: compute_pk_randoms Fisher realisation, mesh %s, box %s, k in [%.3f,%.3f] dk=%.4f lmax=%d, n_b=%d, FKP weights"
                       % (args.workload, "x".join(str(int(g)) for g in grid), "x".join("%g" % b for b in wl["box"]), wl["bins"][0],
                          wl["bins"][1], wl["bins"][2], wl["bins"][3], n_b), "n_b": n_b}
    line = {"impl": "reference", "metric": "pk Fisher MC realisations/s", "value": val, "unit": "realisations/s",
            "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "step_fraction_of_realisation": (wall / args.steps) / T, "seconds_per_realisation": T,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64/f64 (the reference's shipped precision)",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "realisations/s", "cores": arm.threads, "kind": "port",
                             "sample": arm.describe(), "parts_s": {k: v for k, v in parts.items() if k != "check"}},
            "e2e": {"value": val, "unit": "realisations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, mesh):
    return {"workload": "%s: compute_pk_randoms Fisher realisation, mesh %s, box %s, k in [%.3f,%.3f] dk=%.4f lmax=%d, n_b=%d, FKP weights"
            % (name, "x".join(str(int(g)) for g in mesh.grid), "x".join("%g" % b for b in mesh.box), mesh.k_min, mesh.k_max,
               mesh.dk, mesh.lmax, mesh.n_k * mesh.n_l),
            "n_b": mesh.n_k * mesh.n_l, "n_fft_min_per_realisation": n_fft_min(sum(2 * l + 1 for l in range(0, mesh.lmax + 1, 2)), mesh.n_k * mesh.n_l),
            "l2": "inputs>L2 (every map is >= 1 GB, L2 is 126 MB)" if mesh.N * 8 > 2 * 126e6 else "maps fit L2 partially; 512 MB scratch written between steps"}


def run_pk_data(args, wl, mesh, rank, world, local, dev):
    """pk data path (pk/compute_pk_data.py:133-206): TSC painting of galaxies and randoms + q_alpha."""
    import torch
    from spectra_without_windows_b200 import engine
    ctx = engine.Context(mesh, local, ("pk_data",))
    rax_d = [torch.from_numpy(a).to(dev) for a in mesh.rax]
    nbar_r = uniform_nbar_r(rax_d, torch).contiguous()
    ctx.set_fields(nbar_r * ALPHA, nbar_r * ALPHA, nbar_r, SHOT)
    g = torch.Generator(device=dev)
    g.manual_seed(1001 + rank)
    box = torch.tensor(wl["box"], dtype=torch.float64, device=dev)
    bc = torch.tensor(wl["box_center"], dtype=torch.float64, device=dev)
    dpos = (torch.rand((wl["n_gal"], 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * box + bc
    rpos = (torch.rand((wl["n_ran"], 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * box + bc
    dw = torch.ones(wl["n_gal"], dtype=torch.float64, device=dev)
    rw = torch.ones(wl["n_ran"], dtype=torch.float64, device=dev)
    alpha = wl["n_gal"] / wl["n_ran"]
    v = mesh.v_cell

    def step():
        diff = ctx.paint(dpos, dw, wl["box_center"], scale=1.0 / v)
        ctx.paint(rpos, rw, wl["box_center"], scale=-alpha / v, out=diff)
        return ctx.pk_data_q(diff)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        q = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record()
    for _ in range(args.steps):
        diff = ctx.paint(dpos, dw, wl["box_center"], scale=1.0 / v)
        ctx.paint(rpos, rw, wl["box_center"], scale=-alpha / v, out=diff)
    ep1.record()
    torch.cuda.synchronize()
    pms = ep0.elapsed_time(ep1) / args.steps
    npart = wl["n_gal"] + wl["n_ran"]
    print(json.dumps({"metric": "pk data catalogues/s", "value": args.steps / (ms * 1e-3), "unit": "catalogues/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "C1: compute_pk_data (TSC paint of %d galaxies + %d randoms, q_alpha), mesh 128^3, n_b=%d"
                                 % (wl["n_gal"], wl["n_ran"], ctx.n_b)},
                      "paint_ms": pms, "paint_particles_per_s": npart / (pms * 1e-3),
                      "paint_gbs_algorithmic": 32.0 * npart / (pms * 1e-3) / 1e9, "q_finite": bool(torch.isfinite(q).all())}))


def run_bk_data(args, wl, mesh, rank, world, local, dev):
    """Bispectrum data path (bk/compute_bk_data.py:250-406): g-maps of the data, 3-field term, 1-field term over N_mc
    bias simulations whose g-maps are resident in HBM (the reference reloads them from .npz per simulation)."""
    import torch
    from spectra_without_windows_b200 import engine
    ctx = engine.Context(mesh, local, ("bk_fisher", "bk_data"))
    rax_d = [torch.from_numpy(a).to(dev) for a in mesh.rax]
    nbar_r = survey_nbar_r(rax_d, torch).contiguous()
    nbar = (nbar_r * ALPHA).contiguous()
    nbar_A = nbar_r.clone()
    nbar_A[nbar_A == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(nbar, nbar, nbar_A, SHOT)
    sig = torch.sqrt(nbar_A)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    diff = torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64) * torch.sqrt(nbar)
    aA = torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64) * sig
    aB = torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64) * sig
    n_mc = args.n_mc or wl["n_mc"]
    del aA, aB

    def step():
        """One catalogue: its g-maps, the 3-field sums, and the 1-field term over N_mc bias simulations whose Gaussian
        fields (numpy legacy stream, seeds 2..N_mc+1 like bk/compute_bk_data.py:371) and g / g~ maps are produced on the
        device and consumed straight away -- the in-memory hand-off of `mc_driver bk --data` without the Fisher pairs."""
        gd = ctx.bk_gmaps(diff, data=True)
        q = ctx.bk_q3(gd)
        # the fields of the next batch of seeds are drawn on a side stream while this batch is transformed and contracted
        for _seed, a_mc in ctx.field_pipeline(range(2, n_mc + 2), sig):
            g_mc, tg_mc = ctx.bk_gmaps(a_mc)
            ctx.bk_q1_accumulate(gd, g_mc, tg_mc, n_mc, q)
        return q

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    own0, _ = ctx.launch_counts()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        q = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    own1, _ = ctx.launch_counts()
    ctx.profile(True)
    step()
    prof = {k: {"launches": n, "ms": round(m_, 3)} for k, (n, m_) in ctx.profile_report().items()}
    ctx.profile(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # 16 N per transform of the catalogue (M + n_l n_k + ...) + 16 N n_l n_k per bias simulation (its g and g~ maps read once)
    M = sum(2 * l + 1 for l in range(0, mesh.lmax + 1, 2))
    nfft = M + mesh.n_l * mesh.n_k
    b_alg = 16.0 * mesh.N * nfft + 16.0 * mesh.N * mesh.n_l * mesh.n_k * n_mc
    achieved = b_alg * (args.steps / (ms * 1e-3)) / 1e9
    g3 = prof.get("gram3.dmma f64", {}).get("ms")
    flop3 = 3.0 * 2.0 * mesh.n_k * mesh.n_k * mesh.n_l * mesh.n_k * mesh.N * n_mc      # three trilinear Grams per simulation
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": "whole catalogue under the SURVEY 8d convention (16 N x %d transforms + 16 N n_l n_k N_mc map bytes); the 1-field "
                      "term is FP64-tensor work (trilinear Gram), %.3g flop" % (nfft, flop3),
            "gram3": {"ms": g3, "tflops": (flop3 / (g3 * 1e-3) / 1e12) if g3 else None}}
    print(json.dumps({"metric": "bk data catalogues/s", "value": args.steps / (ms * 1e-3), "unit": "catalogues/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "C3: compute_bk_data (g-maps, 3-field term, 1-field term over N_mc=%d bias simulations generated "
                                 "and consumed on the device), mesh 256^3, k in [0,0.11] dk=0.01 l=0,2, %d triangles" % (n_mc, ctx.n_tri),
                                 "l2": "inputs>L2"},
                      "gpu_launches": int(own1 - own0), "clocks": clocks, "roofline": roof, "profile_one_step": prof,
                      "q_finite": bool(torch.isfinite(q).all())}))


def run_bk(args, wl, mesh, rank, world, local, dev):
    """Bispectrum Fisher: one step = one pair of realisations (bk/compute_bk_randoms.py)."""
    import torch
    import torch.distributed as dist
    from spectra_without_windows_b200 import engine
    ctx = engine.Context(mesh, local, ("bk_fisher",))
    if args.fft >= 0:
        ctx.set_fft_backend(args.fft)
    rax_d = [torch.from_numpy(a).to(dev) for a in mesh.rax]
    nbar_r = survey_nbar_r(rax_d, torch).contiguous()
    nbar = (nbar_r * ALPHA).contiguous()
    nbar_A = nbar_r.clone()
    nbar_A[nbar_A == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(nbar, nbar, nbar_A, SHOT)
    sig = torch.sqrt(nbar_A)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    aA = torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64) * sig
    aB = torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64) * sig
    n_b = ctx.n_b_bk
    ctx.bk_delta()
    # K independent pairs in flight (context + stream each): the tensor-pipe-bound Gram of one pair overlaps the
    # HBM-bound passes of the other at the kernel tails
    K = max(1, min(int(args.inflight), 2))
    torch.cuda.synchronize()
    footprint = torch.cuda.memory_allocated(dev)      # at C4 the phi maps of one pair take ~150 GB: one pair at a time
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    K = max(1, min(K, 1 + int((free_b - (24 << 30)) // max(1, footprint))))
    main_stream = torch.cuda.current_stream(dev)
    ctxs, streams = [ctx], [main_stream]
    for k in range(1, K):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            ck = engine.Context(mesh, local, ("bk_fisher",))
            if args.fft >= 0:
                ck.set_fft_backend(args.fft)
            ck.set_fields(nbar, nbar, nbar_A, SHOT)
            ck.bk_delta()
        ctxs.append(ck)
        streams.append(st)
    torch.cuda.synchronize()
    F_acc = [torch.zeros((n_b, n_b), dtype=torch.float64, device=dev) for _ in range(K)]
    maps = [None] * K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        k = i % K
        with torch.cuda.stream(streams[k]):
            F, maps[k] = ctxs[k].bk_fisher_pair(aA, aB, maps[k])
            F_acc[k].add_(F)

    for i in range(max(args.warmup, K)):
        step(i)
    barrier()
    counts0 = [c_.launch_counts() for c_ in ctxs]
    sampler = ClockSampler(local) if rank == 0 else None
    for k in range(K):
        F_acc[k].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(main_stream)
    for st in streams[1:]:
        st.wait_event(e0)
    for i in range(args.steps):
        step(i)
    for st in streams[1:]:
        main_stream.wait_stream(st)
    for k in range(1, K):
        F_acc[0].add_(F_acc[k])
    if world > 1:
        dist.all_reduce(F_acc[0])
    e1.record(main_stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    counts1 = [c_.launch_counts() for c_ in ctxs]
    own0, own1 = sum(c_[0] for c_ in counts0), sum(c_[0] for c_ in counts1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    maps = maps[0]
    # ---- end to end through the public call with HOST buffers: both fields H2D from pinned memory, F D2H, every pair
    e2e = None
    if not args.no_e2e:
        hA, hB = aA.cpu().pin_memory(), aB.cpu().pin_memory()
        dA, dB = torch.empty_like(aA), torch.empty_like(aB)
        F_host = torch.empty((n_b, n_b), dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            dA.copy_(hA, non_blocking=True)
            dB.copy_(hB, non_blocking=True)
            F, maps = ctx.bk_fisher_pair(dA, dB, maps)
            F_host.copy_(F, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        e2e = {"value": 2.0 * world * args.steps / float(tdt.item()), "unit": "realisations/s", "h2d_bytes_per_step": int(2 * mesh.N * 8),
               "d2h_bytes_per_step": int(n_b * n_b * 8), "note": "two fields H2D from pinned memory and F D2H per pair; host RNG not included"}
        del hA, hB, dA, dB
    prof = None
    if rank == 0:
        fft_before = ctx.launch_counts()[1]
        ctx.profile(True)
        ctx.bk_fisher_pair(aA, aB, maps)
        prof = {k: {"launches": n, "ms": round(m_, 3)} for k, (n, m_) in ctx.profile_report().items()}
        ctx.profile(False)
        fft_pair = ctx.launch_counts()[1] - fft_before
        value = 2.0 * world * args.steps / (ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # B_alg = 16 N x N_FFT,min per pair (8 628 transforms for n_k = 11, lmax = 4, 191 triangles) + the Gram flops
        nfft_min = 8628 if (n_b == 573 and mesh.n_k == 11) else int(fft_pair)
        b_alg = 16.0 * mesh.N * nfft_min
        achieved = b_alg * (world * args.steps / (ms * 1e-3)) / world / 1e9
        gram_ms = prof.get("gram.dmma f64", {}).get("ms")
        gram_flop = 2.0 * n_b * n_b * mesh.N
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "kernel": "whole pair under the SURVEY 8d convention: 16 N bytes x %d transforms (the Gram's %.3g flop are extra work, "
                          "not counted as bytes)" % (nfft_min, gram_flop),
                "transforms_performed_per_pair": int(fft_pair),
                "gram": {"flop_full": gram_flop, "flop_computed_upper_trapezoid": gram_flop * (n_b + 1) / (2.0 * n_b), "ms": gram_ms,
                         "tflops_useful_full_equivalent": (gram_flop / (gram_ms * 1e-3) / 1e12) if gram_ms else None,
                         "compare": "bench.py --workload DGEMM times cuBLAS DGEMM and this kernel on the same [n_b x K] shapes"}}
        line = {"metric": "bk Fisher MC realisations/s", "value": value, "unit": "realisations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: compute_bk_randoms Fisher pair (2 realisations per step), mesh %s, k in [%.3f,%.3f] dk=%.3f lmax=%d, %d triangles, n_b=%d"
                           % (args.workload, "x".join(str(int(g)) for g in mesh.grid), mesh.k_min, mesh.k_max, mesh.dk, mesh.lmax, ctx.n_tri, n_b),
                           "l2": "inputs>L2"}, "pairs_in_flight": K,
                "fft_backend": "sm100a-passes" if ctx.fft_backend == 1 else "cufft",
                "gpu_launches": int(own1 - own0), "clocks": clocks, "e2e": e2e, "roofline": roof, "profile_one_pair": prof,
                "gram_flop_per_pair": 2.0 * n_b * n_b * mesh.N}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.ref_arm import bk_pair_sample
                line["cpu_baseline"] = bk_pair_sample(tuple(int(g) for g in mesh.grid), n_b, mesh.n_l, mesh.n_k)
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_dgemm(args, wl, mesh, rank, world, local, dev):
    """FP64 Gram C = A B^T with A, B [n_b x K]: this repo's DMMA kernel (sww_gram_f64, csrc/gram.cu) against cuBLAS DGEMM
    (torch.matmul) on the same buffers -- the measured FP64 matrix throughput of the box, and the comparison SURVEY K6 asks for
    (bk/compute_bk_randoms.py:489-497 is the contraction)."""
    import torch
    from spectra_without_windows_b200 import engine
    ctx = engine.Context(mesh, local, ("ops",))
    n_b, K = 573, 1 << 22
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    A = torch.randn((n_b, K), generator=gen, device=dev, dtype=torch.float64)
    B = torch.randn((n_b, K), generator=gen, device=dev, dtype=torch.float64)
    flop = 2.0 * n_b * n_b * K

    def timeit(fn):
        for _ in range(max(3, args.warmup)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    out = ctx.zeros((n_b, n_b))
    ms_own = timeit(lambda: ctx.gram(A, B, out=out))
    ms_blas = timeit(lambda: torch.matmul(A, B.T))
    ref = torch.matmul(A, B.T)
    err = float((ctx.gram(A, B) - ref).abs().max() / ref.abs().max())
    Bsq = None
    line = {"metric": "FP64 Gram TFLOP/s", "value": flop / (ms_own * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_own, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "DGEMM: C[573 x 573] = A[573 x 2^22] B^T, FP64; inputs 2 x 19 GB > L2"},
            "gram_dmma_kernel": {"ms": ms_own, "tflops": flop / (ms_own * 1e-3) / 1e12,
                                 "note": "computes the upper trapezoid only when A is B's partner in the Fisher assembly; here the full product"},
            "cublas_dgemm": {"ms": ms_blas, "tflops": flop / (ms_blas * 1e-3) / 1e12, "api": "torch.matmul (cuBLAS)"},
            "max_rel_diff": err}
    print(json.dumps(line))


def run_job(args, wl, mesh, ctx, sig, rank, world, local, dev, big):
    """The Monte-Carlo job the reference runs as N_mc SLURM tasks + the fold in compute_pk_data.py
    (pk/compute_pk_randoms.py:135-287 per seed, pk/compute_pk_data.py:227-247): seeds 1..N_mc dealt round-robin to the
    ranks, every field drawn ON THE DEVICE from numpy's legacy stream (seed = rand_it), Fisher realisations with K in
    flight, one all-reduce, rank 0 writes the combined fisher-pk / bias-pk files.  The clock starts when the contexts
    exist and the NCCL communicator is built (the reference's counterpart: interpreter + imports + load_nbar, which it pays
    per seed) and stops when the files are on disk: first field, loop, all-reduce and writes are all inside."""
    import tempfile
    import torch
    import torch.distributed as dist
    from spectra_without_windows_b200 import mc_driver
    N_mc = args.n_mc or wl.get("n_mc_job", 100)
    K = 1 if big else max(1, min(int(args.inflight), 4))
    ctxs, streams = mc_driver.make_inflight_contexts(ctx, K, ("pk_fisher",))
    n_b = ctx.n_b
    outdir = tempfile.mkdtemp(prefix="sww_job_")

    def one(seed, a, c):
        return c.pk_fisher(a)

    # untimed warm-up: one realisation per context (kernel attributes, shell order, RNG scratch)
    mc_driver.pk_device_rng_loop(ctx, [10 ** 6 + rank * 8 + k for k in range(len(ctxs))], sig, one, K, ("pk_fisher",), ctxs, streams)
    if world > 1:
        dist.barrier()                     # also builds the NCCL communicator before the clock starts
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    counts0 = [c_.launch_counts() for c_ in ctxs]
    marks = {}
    t0 = time.perf_counter()
    seeds = mc_driver.shard(range(1, N_mc + 1), rank, world)
    F, q, _ = mc_driver.pk_device_rng_loop(ctx, seeds, sig, one, K, ("pk_fisher",), ctxs, streams, marks)
    t1 = time.perf_counter()
    if world > 1:
        dist.all_reduce(F)
        dist.all_reduce(q)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        np.save(os.path.join(outdir, "fisher-pk_job.npy"), (F / N_mc).cpu().numpy())
        np.save(os.path.join(outdir, "bias-pk_job.npy"), (q / N_mc).cpu().numpy())
    t3 = time.perf_counter()
    tt = torch.tensor([t3 - t0, t1 - t0, t2 - t1, t3 - t2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total, t_loop, t_red, t_write = (float(v) for v in tt.tolist())
    clocks = sampler.stop() if sampler else None
    counts1 = [c_.launch_counts() for c_ in ctxs]
    own = sum(c1[0] - c0[0] for c0, c1 in zip(counts0, counts1))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        M = sum(2 * l + 1 for l in range(0, mesh.lmax + 1, 2))
        b_alg = 16.0 * mesh.N * n_fft_min(M, n_b)
        value = N_mc / total
        achieved = b_alg * value / world / 1e9
        Fh = np.load(os.path.join(outdir, "fisher-pk_job.npy"))
        line = {"metric": "pk Fisher MC realisations/s", "mode": "job", "value": value, "unit": "realisations/s", "n_gpus": world,
                "steps": N_mc, "warmup": len(ctxs), "ms_per_step": total / N_mc * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(args.workload, mesh), N_mc=N_mc,
                               job="seeds 1..N_mc round-robin over the ranks; fields from the device legacy RNG (numpy stream, seed = rand_it)"),
                "realisations_in_flight": len(ctxs), "row_mesh": list(ctx.row_mesh_shape()) if ctx.row_mesh_shape() else None,
                "gpu_launches": int(own), "clocks": clocks,
                "phases_s": {"total": total, "mc_loop": t_loop, "first_field_enqueued": marks.get("first_field_enqueued_s"),
                             "all_reduce": t_red, "write_combined_files": t_write,
                             "realisations_on_busiest_rank": int(np.ceil(N_mc / world))},
                "e2e": {"value": value, "unit": "realisations/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": int((n_b * n_b + n_b) * 8 / max(1, N_mc)),
                        "note": "the job IS the end-to-end path: fields are generated on the device (no host RNG, no H2D), the result leaves once"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "kernel": "whole job under the SURVEY 8d convention (16 N bytes x (2M+1+2 n_b) transforms per realisation), per GPU"},
                "check": {"fisher_trace": float(np.trace(Fh)), "fisher_finite": bool(np.all(np.isfinite(Fh)))}}
        print(json.dumps(line))
    import shutil
    shutil.rmtree(outdir, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--fft", type=int, default=-1, help="FFT backend: 0 cuFFT, 1 hand-written sm_100a passes (default: auto = 1 on power-of-two meshes)")
    ap.add_argument("--inflight", type=int, default=3,
                    help="independent realisations in flight per GPU (one context + stream each): the tail of one "
                         "realisation's kernels overlaps the next one's; 1 = strictly one after the other")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile", action="store_true", help="after the timed region, one extra realisation with CUDA-event profiling per kernel family")
    ap.add_argument("--mode", default="step", choices=["step", "job"],
                    help="job: the whole Monte-Carlo job of the workload (N_mc realisations SHARDED over the ranks = strong scaling): "
                         "fields drawn on the device from numpy's legacy stream, accumulation, all-reduce, combined files written")
    ap.add_argument("--n-mc", type=int, default=0, help="job mode: realisations of the job (default: 100 for C2, 200 for C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from spectra_without_windows_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise Exception("bench.py needs a CUDA device (no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    wl = WORKLOADS[args.workload]
    big = int(np.prod(wl["grid"])) >= 1024 ** 3
    if big:
        os.environ["SWW_TIGHT_WS"] = "1"     # 1024^3: size the workspace for the fused chains only
    mesh = engine.Mesh(wl["box"], wl["grid"], wl["box_center"], *wl["bins"])
    if wl.get("dgemm"):
        return run_dgemm(args, wl, mesh, rank, world, local, dev)
    if wl.get("bk_data"):
        return run_bk_data(args, wl, mesh, rank, world, local, dev)
    if wl.get("bk"):
        return run_bk(args, wl, mesh, rank, world, local, dev)
    if wl.get("data"):
        return run_pk_data(args, wl, mesh, rank, world, local, dev)
    ctx = engine.Context(mesh, local, ("pk_fisher",))
    if args.fft >= 0:
        ctx.set_fft_backend(args.fft)
    rax_d = [torch.from_numpy(a).to(dev) for a in mesh.rax]
    nb_fn = uniform_nbar_r if args.workload == "C1F" else survey_nbar_r
    nbar_r = nb_fn(rax_d, torch).contiguous()
    nbar = (nbar_r * ALPHA).contiguous()
    nbar_A = nbar_r.clone()
    nbar_A[nbar_A == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(nbar, nbar, nbar_A, SHOT)
    del rax_d, nbar_r
    n_b = ctx.n_b
    sig = torch.sqrt(nbar_A)
    if args.mode == "job":
        return run_job(args, wl, mesh, ctx, sig, rank, world, local, dev, big)

    # synthetic Gaussian fields a ~ N(0, nbar_A) (pk/compute_pk_randoms.py:139-140), one per realisation;
    # each rank draws its own independent realisations (seed offset by rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    nbuf = 1 if big else 2
    a_dev = [torch.randn(mesh.shape, generator=gen, device=dev, dtype=torch.float64).mul_(sig) for _ in range(nbuf)]
    if big:
        del sig
        sig = None
        args.no_e2e = True
    # K independent realisations in flight: context k works on stream k (realisations are independent; their sums are
    # combined at the end), so the tails of one realisation's kernels overlap the next one's
    K = 1 if big else max(1, min(int(args.inflight), 4))
    # every extra context brings its own workspace and Y_lm cache: only as many as the device memory holds, leaving
    # room for the end-to-end buffers
    torch.cuda.synchronize()
    footprint = torch.cuda.memory_allocated(dev)
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    K = max(1, min(K, 1 + int((free_b - (12 << 30)) // max(1, footprint))))
    main_stream = torch.cuda.current_stream(dev)
    ctxs, streams = [ctx], [main_stream]
    for k in range(1, K):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            ck = engine.Context(mesh, local, ("pk_fisher",))
            if args.fft >= 0:
                ck.set_fft_backend(args.fft)
            ck.set_fields(nbar, nbar, nbar_A, SHOT)
        ctxs.append(ck)
        streams.append(st)
    torch.cuda.synchronize()
    F_acc = [torch.zeros((n_b, n_b), dtype=torch.float64, device=dev) for _ in range(K)]
    q_acc = [torch.zeros((n_b,), dtype=torch.float64, device=dev) for _ in range(K)]
    F_i = [torch.empty_like(F_acc[0]) for _ in range(K)]
    q_i = [torch.empty_like(q_acc[0]) for _ in range(K)]
    flush = None
    if mesh.N * 8 <= 2 * 126e6:
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        k = i % K
        with torch.cuda.stream(streams[k]):
            ctxs[k].pk_fisher(a_dev[i % nbuf], F_i[k], q_i[k])
            F_acc[k].add_(F_i[k])
            q_acc[k].add_(q_i[k])
            if flush is not None:
                flush.fill_(i & 255)

    def fork(ev):
        ev.record(main_stream)
        for st in streams[1:]:
            st.wait_event(ev)

    def join():
        for st in streams[1:]:
            main_stream.wait_stream(st)

    for i in range(args.warmup):
        step(i)
    barrier()
    counts0 = [c_.launch_counts() for c_ in ctxs]
    sampler = ClockSampler(local) if rank == 0 else None
    for k in range(K):
        F_acc[k].zero_()
        q_acc[k].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    fork(e0)
    for i in range(args.steps):
        step(i)
    join()
    for k in range(1, K):
        F_acc[0].add_(F_acc[k])
        q_acc[0].add_(q_acc[k])
    if world > 1:   # the only exchange of the path: partial Fisher / bias sums (pk/compute_pk_data.py:227-233)
        dist.all_reduce(F_acc[0])
        dist.all_reduce(q_acc[0])
    e1.record(main_stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    counts1 = [c_.launch_counts() for c_ in ctxs]
    own0, fft0 = (sum(c_[j] for c_ in counts0) for j in (0, 1))
    own1, fft1 = (sum(c_[j] for c_ in counts1) for j in (0, 1))
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * args.steps / (ms * 1e-3)
    F_iK, q_iK = F_i, q_i
    F_i, q_i = F_iK[0], q_iK[0]

    # ---- end-to-end through the public API with HOST buffers (pinned): H2D of the field, D2H of F and q-bar
    e2e = None
    if not args.no_e2e:
        a_host = [torch.empty(mesh.shape, dtype=torch.float64).pin_memory() for _ in range(2)]
        for h, d in zip(a_host, a_dev):
            h.copy_(d)
        F_host = [torch.empty((n_b, n_b), dtype=torch.float64).pin_memory() for _ in range(K)]
        q_host = [torch.empty((n_b,), dtype=torch.float64).pin_memory() for _ in range(K)]
        copy_stream = torch.cuda.Stream(device=dev)
        B = 2 * K                                  # device field buffers: K in use, K being filled
        a_e2e = list(a_dev) + [torch.empty_like(a_dev[0]) for _ in range(B - len(a_dev))]
        ready = [torch.cuda.Event() for _ in range(B)]
        freed = [torch.cuda.Event() for _ in range(B)]

        def e2e_run(n):
            """Every realisation: H2D of its field from pinned memory (copy stream, K ahead), the Fisher realisation on
            context i % K, D2H of F and q-bar."""
            for j in range(B):
                freed[j].record(main_stream)

            def issue_copy(i):
                j = i % B
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[j])
                    a_e2e[j].copy_(a_host[i % 2], non_blocking=True)
                    ready[j].record(copy_stream)

            for i in range(min(K, n)):
                issue_copy(i)
            for i in range(n):
                j, k = i % B, i % K
                if i + K < n:
                    issue_copy(i + K)
                st = streams[k]
                st.wait_event(ready[j])
                with torch.cuda.stream(st):
                    ctxs[k].pk_fisher(a_e2e[j], F_iK[k], q_iK[k])
                    freed[j].record(st)
                    F_host[k].copy_(F_iK[k], non_blocking=True)
                    q_host[k].copy_(q_iK[k], non_blocking=True)
            torch.cuda.synchronize()

        e2e_run(1)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        dt = float(tdt.item())
        # the same loop with the field drawn ON THE DEVICE from numpy's legacy stream (no host RNG, no H2D)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, (seed, a_f) in enumerate(ctx.field_pipeline([1000 * rank + i + 1 for i in range(args.steps)], sig, consumers=streams)):
            k = i % K
            with torch.cuda.stream(streams[k]):
                ctxs[k].pk_fisher(a_f, F_iK[k], q_iK[k])
                F_host[k].copy_(F_iK[k], non_blocking=True)
                q_host[k].copy_(q_iK[k], non_blocking=True)
        torch.cuda.synchronize()
        dt_rng = time.perf_counter() - t0
        trng = torch.tensor([dt_rng], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(trng, op=dist.ReduceOp.MAX)
        dt_rng = float(trng.item())
        e2e = {"value": world * args.steps / dt, "unit": "realisations/s", "h2d_bytes_per_step": int(mesh.N * 8),
               "with_device_legacy_rng": {"value": world * args.steps / dt_rng, "unit": "realisations/s",
                                          "note": "numpy-legacy Gaussian fields generated on the device (sww_grf_legacy_batch, four seeds side by side on a side stream, "
                                                  "overlapped) + Fisher realisations; short runs pay the first, un-overlapped batch (the N_mc = 100 product "
                                                  "run reaches 4.1/s, DESIGN.md); the reference draws a field on the host in ~5.4 s per 512^3 per core"},
               "d2h_bytes_per_step": int((n_b * n_b + n_b) * 8), "note": "field H2D from pinned memory (copy stream, %d realisations ahead), F and q-bar D2H every step; host RNG not included" % K}

    prof = None
    if rank == 0 and (args.profile or ctx.fft_backend == 1):
        ctx.profile(True)
        ctx.pk_fisher(a_dev[0], F_i, q_i)
        prof = {k: {"launches": n, "ms": round(ms_, 3)} for k, (n, ms_) in ctx.profile_report().items()}
        ctx.profile(False)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        M = sum(2 * l + 1 for l in range(0, mesh.lmax + 1, 2))
        nfft = n_fft_min(M, n_b)
        b_alg = 16.0 * mesh.N * nfft            # bytes per realisation (SURVEY.md section 8d convention)
        achieved = b_alg * (args.steps / (ms * 1e-3)) / 1e9   # per GPU
        step_roof = {"achieved": achieved, "frac": achieved / peak, "unit": "GB/s",
                     "algorithmic_bytes_per_realisation": b_alg,
                     "definition": "16 N bytes per 3-D transform x (2M+1+2n_b) transforms per realisation (SURVEY.md 8d) / measured step time"}
        rm = ctx.row_mesh_shape()
        if rm:
            # the 2 n_b row transforms run on the row mesh (csrc/rowmesh.cu): bytes of the transforms as executed, so that the
            # fraction stays a fraction (the survey convention charges the full mesh for them and can exceed 1)
            b_exec = 16.0 * (mesh.N * (2 * M + 1) + float(np.prod(rm)) * 2 * n_b)
            achieved_exec = b_exec * (args.steps / (ms * 1e-3)) / 1e9
            step_roof.update({"frac_survey_convention": achieved / peak, "achieved_survey_convention": achieved,
                              "algorithmic_bytes_as_executed": b_exec, "achieved": achieved_exec, "frac": achieved_exec / peak,
                              "definition": "16 bytes per cell and 3-D transform: 2M+1 transforms on the mesh + 2 n_b on the row mesh %s / measured "
                                            "step time (frac_survey_convention: all of them charged at the full mesh, SURVEY.md 8d)" % (list(rm),)})
            achieved = achieved_exec
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "kernel": "whole realisation step", "step": step_roof}
        zkey = "fft.Z c2r*w*r2c fused"
        if prof is not None and zkey in prof:
            # dominant kernel: the fused z pass (c2r * nW * r2c in shared memory).  Algorithmic bytes per launch =
            # the weight map (8 N) + the complex input rows of the shell's box + the complex output rows of the k_max box.
            rmesh = ctx.row_mesh_shape()          # the rows run on the coarse row mesh when there is one (csrc/rowmesh.cu)
            nx, ny, nz = rmesh if rmesh else (int(v) for v in mesh.grid)
            kFz = 2.0 * np.pi / mesh.box[2]
            kz_of = lambda kmax: min(int(np.floor(np.nextafter(kmax / kFz, 0))) + 1, nz // 2 + 1)
            kz_out = kz_of(mesh.k_hi[-1])
            kz_in_mean = float(np.mean([kz_of(k) for k in mesh.k_hi]))
            rows_per_launch = max(1, round(n_b / prof[zkey]["launches"]))   # the n_l multipoles of one k bin share a launch
            bytes_launch = 8.0 * nx * ny * nz + rows_per_launch * 16.0 * nx * ny * (kz_in_mean + kz_out)
            dur = prof[zkey]["ms"] / prof[zkey]["launches"] * 1e-3
            traffic = None
            try:
                if args.workload != "C2":
                    raise KeyError("the ncu capture was taken on C2")
                tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
                if tj.get("rows_per_launch", 1) != rows_per_launch:
                    raise KeyError("capture taken with a different batching")
                traffic = tj["dram_bytes_per_launch"]
            except Exception:
                pass
            roof["kernel"] = ("whole realisation step = %d fused 5-pass chains (I1,I2,Z,F2,F3) + 2M+1 forward chains; BASELINE.md section 3 "
                              "defines the metric's %% of HBM roofline on this unit" % prof[zkey]["launches"])
            roof["traffic"] = traffic
            roof["dominant_kernel"] = {
                "name": "%s (fused z pass: c2r * nW * r2c of the %d Fisher rows of one k bin)"
                        % ("pass_z16p_fused_kernel<NBI,true,3>" if nz == 512 and kz_out <= 128 else
                           ("pass_zw_kernel<H,1,1,0>" if nz & (nz - 1) == 0 and nz >= 128 else "gpass_z_kernel<1,1>"), rows_per_launch),
                "launches_per_realisation": prof[zkey]["launches"], "avg_launch_ms": dur * 1e3,
                "share_of_step": prof[zkey]["ms"] / (ms / args.steps),
                "algorithmic_bytes_per_launch": bytes_launch, "achieved_gbs": bytes_launch / dur / 1e9,
                "frac_of_hbm_peak": bytes_launch / dur / 1e9 / peak, "dram_traffic_bytes_per_launch_ncu": traffic,
                "row_mesh": list(rmesh) if rmesh else None,
                "note": "FP64-pipe bound (ncu: 56 % of the pipe, 64 DFMA/clk/SM; two half-length complex FFTs per row), so its HBM fraction is below the step's; see profiles/ for the ncu captures",
                "timing": "CUDA events on the launching stream around every launch of one extra realisation after the timed region"}
        line = {"metric": "pk Fisher MC realisations/s", "value": value, "unit": "realisations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args.workload, mesh), "realisations_in_flight": K,
                "row_mesh": list(ctx.row_mesh_shape()) if ctx.row_mesh_shape() else None,
                "fft_backend": "sm100a-passes" if ctx.fft_backend == 1 else "cufft",
                "gpu_launches": int((own1 - own0) + (fft1 - fft0)) if ctx.fft_backend == 0 else int(own1 - own0),
                "own_kernel_launches": int(own1 - own0), "library_fft_calls": int(fft1 - fft0) if ctx.fft_backend == 0 else 0,
                "clocks": clocks, "e2e": e2e, "roofline": roof}
        if prof is not None:
            line["profile_one_realisation"] = prof
        if world == 1 and not args.no_cpu_baseline:
            try:
                arm = reference_arm(wl, args.workload)
                arm.prologue()
                dt = arm.step()[0] + arm.step()[0]
                T, parts = arm.T_real(), arm.parts()
                arm.close()
                line["cpu_baseline"] = {"value": 1.0 / T, "unit": "realisations/s", "cores": arm.threads, "kind": "port",
                                        "sample": arm.describe() + "; one loop-1 row + one loop-2 row step, %.1f s" % dt,
                                        "parts_s": {k: v for k, v in parts.items() if k != "check"}}
            except Exception as ex:   # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
