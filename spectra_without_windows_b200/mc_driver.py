"""Monte-Carlo Fisher driver: all N_mc realisations of a .param file on 1..8 GPUs.

The reference runs one OS process per seed under SLURM and reduces through files
(slurm/pk_randoms_boss_nC.slurm:24-35, pk/compute_pk_data.py:227-233).  Here the seeds are dealt
round-robin to the ranks of one torch.distributed job (one process per GPU), every rank accumulates
its partial sums on the device, and ONE all-reduce of n_b^2 + n_b doubles ends the run; rank 0 writes
the combined `fisher-pk_*` / `bias-pk_*` (or `fisher-bk_*`) files the data scripts read.  There is no
data-path collective: realisations are independent.

    torchrun --nproc-per-node 8 -m spectra_without_windows_b200.mc_driver pk <paramfile> [--per-seed]
    torchrun --nproc-per-node 8 -m spectra_without_windows_b200.mc_driver bk <paramfile> [--data 1,2,..] [--no-bias-maps]

`--data s1,s2,..` (bispectrum): in-memory hand-off to the data step -- the listed catalogues' g-maps stay on the device,
every Monte-Carlo seed's g / g~ maps feed their 1-field sums straight from the Fisher pair, one all-reduce of the sums, and
rank 0 writes the `bk_*.txt` files compute_bk_data.py would write.  The `*_bk_bias_alpha_*.npz` files of the reference's
contract are written asynchronously (complex64, like the reference) unless `--no-bias-maps`.

`--inflight K` (default 3, P_l(k) with device-generated fields): K realisations are kept in flight per GPU, each on
its own context and stream, as far as the device memory holds K workspaces.
`--per-seed` also writes the reference's per-seed checkpoint files, and seeds whose files already
exist are skipped (the reference's idempotent re-entry, pk/compute_pk_randoms.py:121-127).
The Gaussian fields follow numpy's legacy stream with the reference's seeds
(pk/compute_pk_randoms.py:139-140).  By default they are generated ON THE DEVICE (`sww_grf_legacy`:
bit-exact uniforms and accept/reject decisions, Gaussian values within 2 ulp of numpy's; 0.2 s per
512^3 field against 5.4 s on a host core); `--host-rng` draws them with numpy on a small pool of host
threads that runs ahead of the GPU instead (bit-identical inputs to the reference).
"""
from __future__ import annotations

import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def shard(items, rank, world):
    """Round-robin deal of work items to ranks."""
    return list(items)[rank::world]


def reduce_sums(parts, world_ok, dist=None):
    """All-reduce (sum) a list of tensors in place when running distributed."""
    if world_ok and dist is not None:
        for p in parts:
            dist.all_reduce(p)
    return parts


def run_mc(seeds, realisation_fn, zeros_fn, rank=0, world=1, dist=None, prefetch_fn=None, workers=2):
    """Generic sharded Monte-Carlo accumulation.

    seeds           : all work items (e.g. rand_it = 1..N_mc, or pair indices for the bispectrum)
    realisation_fn  : (seed, prefetched) -> tuple of tensors to accumulate
    zeros_fn        : () -> tuple of zero tensors with the shapes of the sums
    prefetch_fn     : seed -> host object (e.g. the numpy Gaussian field), run ahead on host threads
    Returns (sums after all-reduce, number of items this rank processed).
    """
    mine = shard(seeds, rank, world)
    sums = zeros_fn()
    pool = ThreadPoolExecutor(max_workers=max(1, workers)) if prefetch_fn else None
    futs = {}
    if pool:
        for s in mine[:workers + 1]:
            futs[s] = pool.submit(prefetch_fn, s)
    for i, s in enumerate(mine):
        pre = None
        if pool:
            pre = futs.pop(s).result()
            nxt = i + workers + 1
            if nxt < len(mine):
                futs[mine[nxt]] = pool.submit(prefetch_fn, mine[nxt])
        out = realisation_fn(s, pre)
        for acc, o in zip(sums, out):
            acc += o
    if pool:
        pool.shutdown()
    reduce_sums(sums, world > 1, dist)
    return sums, len(mine)


def make_inflight_contexts(ctx, inflight, pipe):
    """K contexts on K streams (the first is `ctx` itself), as many as the device memory holds workspaces for."""
    import torch
    torch.cuda.synchronize()
    footprint = torch.cuda.memory_allocated()
    free_b, _ = torch.cuda.mem_get_info()
    K = max(1, min(inflight, 4, 1 + int((free_b - (12 << 30)) // max(1, footprint))))
    ctxs, streams = [ctx], [ctx.stream]
    for k in range(1, K):
        st = torch.cuda.Stream()
        ctxs.append(ctx.clone_for_stream(st, pipe))
        streams.append(st)
    torch.cuda.synchronize()
    return ctxs, streams


def pk_device_rng_loop(ctx, seeds, sig_dev, one, inflight, pipe, ctxs=None, streams=None, marks=None):
    """The P_l(k) Monte-Carlo loop of one rank with device-generated fields: batch j+1 is generated on a side stream while
    the Fisher kernels of batch j run; K contexts on K streams keep K realisations in flight.  `one(seed, field, ctx)`
    -> (F, q).  Returns (sum F, sum q, time the loop started); `marks` (dict) receives host time stamps of the phases."""
    import time
    import torch
    if ctxs is None:
        ctxs, streams = make_inflight_contexts(ctx, inflight, pipe)
    K = len(ctxs)
    n_b = ctx.n_b
    Fk = [ctx.zeros((n_b, n_b)) for _ in range(K)]
    qk = [ctx.zeros((n_b,)) for _ in range(K)]
    torch.cuda.synchronize()
    t_loop = time.time()      # the extra contexts are set-up, like the first one
    for i, (seed, a) in enumerate(ctx.field_pipeline(seeds, sig_dev, consumers=streams)):
        k = i % K
        if i == 0 and marks is not None:
            marks["first_field_enqueued_s"] = time.time() - t_loop
        with torch.cuda.stream(streams[k]):
            Fi, qi = one(seed, a, ctxs[k])
            Fk[k] += Fi
            qk[k] += qi
    torch.cuda.synchronize()
    if marks is not None:
        marks["loop_s"] = time.time() - t_loop
    F, q = Fk[0], qk[0]
    for k in range(1, K):
        F += Fk[k]
        q += qk[k]
    return F, q, t_loop


def main(argv):
    import torch
    import torch.distributed as dist
    from . import drivers

    if len(argv) < 3 or argv[1] not in ("pk", "bk"):
        raise Exception("usage: mc_driver.py <pk|bk> <paramfile> [--per-seed]")
    spec, paramfile = argv[1], argv[2]
    per_seed = "--per-seed" in argv
    host_rng = "--host-rng" in argv
    inflight = int(argv[argv.index("--inflight") + 1]) if "--inflight" in argv else 3
    data_sims = [int(v) for v in argv[argv.index("--data") + 1].split(",")] if "--data" in argv else []
    write_maps = "--no-bias-maps" not in argv
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = drivers.Params(paramfile, spec)
    P.unsupported()
    for d in (P.outdir, P.mcdir):
        os.makedirs(d, exist_ok=True)
    tag = P.ktag()
    pipe = (("pk_fisher",) if P.wtype == 0 else ("pk_ml",)) if spec == "pk" else (("bk_fisher", "bk_data") if data_sims else ("bk_fisher",))
    ctx, mesh, nbar_A, alpha_ran, shot_fac, _ = drivers._setup(P, 1, pipe, paint_data=False)
    if P.wtype == 1:
        P.load_fiducial_pk(ctx)
    sig = np.sqrt(nbar_A)

    sig_dev = None if host_rng else ctx.dev(sig)
    import time
    if world > 1:
        dist.barrier()       # also builds the NCCL communicator now, not inside the final all-reduce
    torch.cuda.synchronize()
    t_loop = time.time()

    def grf(seed):
        # numpy's legacy global stream is not thread-safe: a private RandomState gives the same values
        # as np.random.seed(seed); np.random.normal(...)  (pk/compute_pk_randoms.py:139-140)
        return np.random.RandomState(seed).normal(loc=0.0, scale=sig, size=sig.shape)

    if spec == "pk":
        n_b = ctx.n_b
        fish_name = lambda r: P.mcdir + "%s%d_%s_pk_fish_a_%s.npy" % (P.string, r, P.weight_type, tag)
        bias_name = lambda r: P.mcdir + "%s%d_%s_pk_q-bar_a_%s.npy" % (P.string, r, P.weight_type, tag)

        def one(seed, a, c=None):
            c = c or ctx
            if per_seed and os.path.exists(fish_name(seed)) and os.path.exists(bias_name(seed)):
                return c.dev(np.load(fish_name(seed))), c.dev(np.load(bias_name(seed)))
            if P.wtype == 1:
                F, q, _ = c.pk_fisher_ml(a, max_it=30, rel_tol=1e-6)
            else:
                F, q = c.pk_fisher(a)
            if per_seed:
                np.save(fish_name(seed), F.cpu().numpy())
                np.save(bias_name(seed), q.cpu().numpy())
            return F, q

        if host_rng:
            (F, q), n = run_mc(range(1, P.N_mc + 1), one, lambda: (ctx.zeros((n_b, n_b)), ctx.zeros((n_b,))), rank, world,
                               dist if world > 1 else None, prefetch_fn=grf)
        else:
            F, q, t_loop = pk_device_rng_loop(ctx, shard(range(1, P.N_mc + 1), rank, world), sig_dev, one, inflight if P.wtype == 0 else 1, pipe)
            reduce_sums([F, q], world > 1, dist if world > 1 else None)
        if rank == 0:
            ctx.mc_finalize(F, q, P.N_mc)
            np.save(P.outdir + "fisher-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag), F.cpu().numpy())
            np.save(P.outdir + "bias-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag), q.cpu().numpy())
    else:
        n_b = ctx.n_b_bk
        fish_name = lambda r: P.mcdir + "%s%d_%s_bk_fish_alpha_beta_%s.npy" % (P.string, r, P.weight_type, tag)
        bias_name = lambda ii: P.mcdir + "%s%d_%s_bk_bias_alpha_%s.npz" % (P.string, ii, P.weight_type, tag)
        maps = None
        # In-memory hand-off of the g / g~ maps to the data step (bk/compute_bk_randoms.py:273-274 ->
        # bk/compute_bk_data.py:370-406): with `--data s1,s2,..` the catalogues' own maps are computed first and stay on
        # the device, and every Monte-Carlo seed's maps go straight from the Fisher pair into the 1-field sums -- they
        # never leave HBM.  The `.npz` contract files are still written (asynchronously, complex64) unless --no-bias-maps.
        held = {}
        if data_sims:
            for s in data_sims:
                dat, rnd, alpha_s, shot_s, bc_s = drivers._catalogues(P, s)
                nbar_s, mask_s, nA_s, diff_s = drivers._fields(P, ctx, dat, rnd, bc_s, alpha_s, True)
                ctx.set_fields(nbar_s, mask_s, nA_s, shot_s)
                g_s = ctx.bk_gmaps(diff_s, data=True)
                held[s] = (g_s, ctx.bk_q3(g_s), ctx.zeros((ctx.n_tri, mesh.n_l)))
            dat, rnd, alpha_1, shot_1, bc_1 = drivers._catalogues(P, 1)          # back to the Monte-Carlo fields
            nbar_1, mask_1, nA_1, _ = drivers._fields(P, ctx, dat, rnd, bc_1, alpha_1, False)
            ctx.set_fields(nbar_1, mask_1, nA_1, shot_1)
            torch.cuda.synchronize()
            t_loop = time.time()
        writer = drivers.GmapWriter(ctx) if write_maps else None

        def one(pair, ab):
            nonlocal maps
            aA = ab[0] if host_rng else ctx.grf_legacy(2 * pair, sig_dev)
            aB = ab[1] if host_rng else ctx.grf_legacy(2 * pair + 1, sig_dev)
            F, maps = ctx.bk_fisher_pair(aA, aB, maps)
            for g_s, _, q1_s in held.values():
                ctx.bk_q1_accumulate(g_s, maps[0], maps[1], P.N_mc, q1_s)
                ctx.bk_q1_accumulate(g_s, maps[2], maps[3], P.N_mc, q1_s)
            if writer:
                # the g-maps of every seed feed the 1-field term of compute_bk_data.py (bk/compute_bk_randoms.py:273-274)
                writer.save(bias_name(2 * pair), maps[0], maps[1])
                writer.save(bias_name(2 * pair + 1), maps[2], maps[3])
            if per_seed:
                np.save(fish_name(pair), F.cpu().numpy())
            return (F,)

        (F,), n = run_mc(range(1, P.N_mc // 2 + 1), one, lambda: (ctx.zeros((n_b, n_b)),), rank, world,
                         dist if world > 1 else None,
                         prefetch_fn=(lambda p: (grf(2 * p), grf(2 * p + 1))) if host_rng else None, workers=1)
        if writer:
            writer.close()
        reduce_sums([h[2] for h in held.values()], world > 1, dist if world > 1 else None)
        if rank == 0:
            fish = (F / (P.N_mc // 2)).cpu().numpy()
            np.save(P.outdir + "fisher-bk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag), fish)
            if held:
                # bk/compute_bk_data.py:409-451 for every held catalogue
                Delta = ctx.bk_delta().cpu().numpy()
                nt, n_l = ctx.n_tri, mesh.n_l
                for s, (_, q3, q1) in held.items():
                    q = (q3 + q1 if P.use_qbar else q3).cpu().numpy()
                    q_alpha = np.concatenate([q[:, l_i] / Delta[l_i * nt:(l_i + 1) * nt] for l_i in range(n_l)])
                    p_alpha = np.matmul(np.linalg.inv(fish), q_alpha)
                    if s != -1:
                        name = P.outdir + "bk_%s%d_%s_N%d_%s.txt" % (P.string, s, P.weight_type, P.N_mc, tag)
                    else:
                        name = P.outdir + "bk_%s%s_N%d_%s.txt" % (P.string, P.weight_type, P.N_mc, tag)
                    drivers.write_bk_text(name, P, s, p_alpha, mesh.triangles())
    torch.cuda.synchronize()
    if rank == 0:
        dt = time.time() - t_loop
        print("%s Monte-Carlo loop: %d realisations on %d GPU(s) in %.2f s (%.3f realisations/s, fields %s)"
              % (spec, P.N_mc, world, dt, P.N_mc / dt, "from numpy on the host" if host_rng else "generated on the device"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv)
