"""The four driver scripts of the reference, re-hosted on the device pipelines.

Same CLI (`python <script> <int> <paramfile>`), same `.param` format, same file names and payloads,
same stdout milestones and the same bare `Exception`s / silent early exits as
pk/compute_pk_randoms.py, pk/compute_pk_data.py, bk/compute_bk_randoms.py, bk/compute_bk_data.py.
The bodies between "LOAD DATA" and "SAVE" are replaced by calls into engine.Context.
"""
from __future__ import annotations

import configparser
import os
import shutil
import sys
import time

import numpy as np

from . import engine
from .opt_utilities import box_center_of, load_data, load_nbar, load_randoms


class Params:
    """Everything the scripts read from the .param file (paramfiles/boss_nC.param:1-74)."""

    def __init__(self, paramfile, spec):
        config = configparser.ConfigParser(interpolation=None, converters={"list": lambda x: [i.strip() for i in x.split(",")]})
        if not os.path.exists(paramfile):
            raise Exception("Parameter file does not exist!")
        config.read(paramfile)
        self.config = config
        self.string = config["sample"]["type"]
        self.weight_type = str(config["settings"]["weights"])
        if self.weight_type == "FKP":
            self.wtype = 0
        elif self.weight_type == "ML":
            self.wtype = 1
        else:
            raise Exception("weights must be 'FKP' or 'ML'")
        sec = "%s-binning" % spec
        self.k_min, self.k_max, self.dk = float(config[sec]["k_min"]), float(config[sec]["k_max"]), float(config[sec]["dk"])
        assert self.k_max > self.k_min
        assert self.dk > 0
        assert self.k_min >= 0
        self.lmax = int(config[sec]["lmax"])
        assert (self.lmax // 2) * 2 == self.lmax, "l-max must be even!"
        self.outdir = str(config["directories"]["output"])
        self.mcdir = str(config["directories"]["monte_carlo"])
        self.tmpdir = str(config["directories"]["temporary"])
        self.ZMIN, self.ZMAX = float(config["sample"]["z_min"]), float(config["sample"]["z_max"])
        self.h_fid, self.OmegaM_fid = float(config["parameters"]["h_fid"]), float(config["parameters"]["OmegaM_fid"])
        self.boxsize_grid = np.array(config.getlist("sample", "box"), dtype=float)
        self.grid_3d = np.array(config.getlist(sec, "grid"), dtype=int)
        self.N_mc = int(config["settings"]["N_mc"])
        self.include_pix = config.getboolean("settings", "include_pix")
        self.rand_nbar = config.getboolean("settings", "rand_nbar")
        self.use_qbar = config.getboolean("settings", "use_qbar", fallback=True)
        self.spec = spec
        if self.wtype == 1:
            # Fiducial power spectrum input (for ML weights), pk/compute_pk_randoms.py:69-71
            self.pk_input_file = str(config["settings"]["fiducial_pk"])

    def unsupported(self):
        """ML weights are built for the P_l(k) scripts.  The bispectrum scripts' ML branch cannot run in the reference
        either (`np.maximum([lmax_pk,lmax])` raises TypeError at bk/compute_bk_randoms.py:173 and
        bk/compute_bk_data.py:170), so there is nothing to be in parity with: fail loudly."""
        if self.wtype == 1 and self.spec != "pk":
            raise Exception("weights = ML is not implemented for the bispectrum scripts on the B200 path yet; use weights = FKP")

    def load_fiducial_pk(self, ctx):
        """pk/compute_pk_randoms.py:188-193: table k | P_0 | P_2 | ... -> P_l(|k|) maps on the device."""
        ctx.set_fiducial_pk(np.loadtxt(self.pk_input_file))

    def ktag(self):
        return "k%.3f_%.3f_%.3f_l%d" % (self.k_min, self.k_max, self.dk, self.lmax)

    def summary(self, extra):
        print("\n###################### PARAMETERS ######################\n")
        print("Data-type: %s" % self.string)
        for line in extra:
            print(line)
        print("Weight-Type: %s" % self.weight_type)
        print("n-bar: from randoms" if self.rand_nbar else "n-bar: from mask")
        print("Forward model pixellation: %d" % self.include_pix)
        print("\nk-min: %.3f" % self.k_min)
        print("k-max: %.3f" % self.k_max)
        print("dk: %.3f" % self.dk)
        print("l-max: %d" % self.lmax)
        print("\nFiducial h = %.3f" % self.h_fid)
        print("Fiducial Omega_m = %.3f" % self.OmegaM_fid)
        print("\nOutput Directory: %s" % self.outdir)
        print("\nBias / Fisher Directory: %s" % self.mcdir)
        print("\n########################################################")


def _catalogues(P: Params, sim_no):
    """Catalogues of one simulation -> alpha, shot factor, BoxCenter (pk/compute_pk_randoms.py:142-153)."""
    from .ingest import Cosmology
    cosmo_coord = Cosmology(h=P.h_fid).match(Omega0_m=P.OmegaM_fid)   # pk/compute_pk_randoms.py:132
    data = load_data(sim_no, P.config, cosmo_coord, fkp_weights=False)
    rand = load_randoms(P.config, cosmo_coord, fkp_weights=False)
    alpha_ran = (data["WEIGHT"].sum() / rand["WEIGHT"].sum()).compute()
    shot_fac = ((data["WEIGHT"] ** 2.0).mean().compute() + alpha_ran * (rand["WEIGHT"] ** 2.0).mean().compute()) / rand["WEIGHT"].mean().compute()
    return data, rand, alpha_ran, shot_fac, box_center_of(rand)


def _fields(P: Params, ctx, data, rand, bc, alpha_ran, paint_data):
    """n(r) maps of one simulation (and its painted field): pk/compute_pk_randoms.py:155-182."""
    nbar_A = load_nbar(P.config, P.spec, 1.0)
    nbar_A[nbar_A == 0] = nbar_A[nbar_A != 0].min() * 0.01
    print("Loading nbar from mask")
    nbar_mask = load_nbar(P.config, P.spec, alpha_ran)
    diff = None
    nbar = nbar_mask
    if paint_data or P.rand_nbar:
        d = ctx.grid_data(data["Position"].v, data["WEIGHT"].v, rand["Position"].v, rand["WEIGHT"].v, bc, return_randoms=P.rand_nbar)
        if P.rand_nbar:
            print("Loading nbar from random particles")
            diff, nbar = d
        else:
            diff = d
    return nbar, nbar_mask, nbar_A, diff


def _setup(P: Params, sim_no, pipelines, paint_data):
    """Catalogues -> alpha, shot factor, BoxCenter; n(r) maps; context with fields set.
    Restates pk/compute_pk_randoms.py:142-182 (identical blocks in the other scripts)."""
    data, rand, alpha_ran, shot_fac, bc = _catalogues(P, sim_no)
    mesh = engine.Mesh(P.boxsize_grid, P.grid_3d, bc, P.k_min, P.k_max, P.dk, P.lmax)
    ctx = engine.Context(mesh, int(os.environ.get("LOCAL_RANK", "0")), pipelines)
    nbar, nbar_mask, nbar_A, diff = _fields(P, ctx, data, rand, bc, alpha_ran, paint_data)
    ctx.set_fields(nbar, nbar_mask, nbar_A, shot_fac)
    ctx.set_include_pix(P.include_pix)
    return ctx, mesh, nbar_A, alpha_ran, shot_fac, diff


def _argv(argv, msg):
    if len(argv) != 3:
        raise Exception(msg)
    return int(argv[1]), str(argv[2])


# ------------------------------------------------------------------------------ pk randoms
def pk_randoms_main(argv):
    """pk/compute_pk_randoms.py: q-bar_alpha and F_ab of one Gaussian realisation."""
    rand_it, paramfile = _argv(argv, "Need to specify random iteration and parameter file!")
    P = Params(paramfile, "pk")
    if rand_it > P.N_mc:
        raise Exception("Simulation number cannot be greater than number of bias sims!")
    if not os.path.exists(P.outdir):
        os.makedirs(P.outdir)
    if not os.path.exists(P.mcdir):
        os.makedirs(P.mcdir)
    P.summary(["Random iteration: %d" % rand_it])
    init = time.time()
    tag = P.ktag()
    bias_file_name = P.mcdir + "%s%d_%s_pk_q-bar_a_%s.npy" % (P.string, rand_it, P.weight_type, tag)
    fish_file_name = P.mcdir + "%s%d_%s_pk_fish_a_%s.npy" % (P.string, rand_it, P.weight_type, tag)
    combined_bias = P.outdir + "bias-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag)
    combined_fish = P.outdir + "fisher-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag)
    if os.path.exists(bias_file_name) and os.path.exists(fish_file_name):
        print("Simulation already completed; exiting!\n")
        sys.exit()
    if os.path.exists(combined_bias) and os.path.exists(combined_fish):
        print("Full Fisher matrices already completed; exiting!\n")
        sys.exit()
    P.unsupported()
    print("\n## Loading %s random iteration %d with %s weights" % (P.string, rand_it, P.weight_type))
    print("\nLoading data")
    ctx, mesh, nbar_A, alpha_ran, shot_fac, _ = _setup(P, 1, ("pk_fisher",) if P.wtype == 0 else ("pk_ml",), paint_data=False)
    print("Data-to-random ratio: %.3f, shot-noise factor: %.3f" % (alpha_ran, shot_fac))
    # Gaussian sample (numpy legacy stream, pk/compute_pk_randoms.py:139-140)
    np.random.seed(rand_it)
    a = np.random.normal(loc=0.0, scale=np.sqrt(nbar_A), size=nbar_A.shape)
    print("\n## Computing Fisher and bias term on the GPU (%d bins)" % ctx.n_b)
    if P.wtype == 0:
        F, qbar = ctx.pk_fisher(a)
    else:
        # ML weights: C^-1 by preconditioned conjugate gradients on every derivative map
        # (pk/compute_pk_randoms.py:211,240 -- rel_tol=1e-6, max_it=30)
        P.load_fiducial_pk(ctx)
        F, qbar, its = ctx.pk_fisher_ml(a, max_it=30, rel_tol=1e-6)
        print("%d conjugate-gradient solves, %d iterations in total" % (ctx.n_b + 1, its))
    np.save(bias_file_name, qbar.cpu().numpy())
    np.save(fish_file_name, F.cpu().numpy())
    duration = time.time() - init
    print("\n## Saved output to %s and %s. Exiting after %d seconds (%d minutes)\n" % (bias_file_name, fish_file_name, duration, duration // 60))
    sys.exit()


# ------------------------------------------------------------------------------ pk data
def pk_data_main(argv):
    """pk/compute_pk_data.py: q_alpha of a catalogue, combination with the MC Fisher/bias, P_l(k) text."""
    sim_no, paramfile = _argv(argv, "Need to specify simulation number and parameter file")
    P = Params(paramfile, "pk")
    extra = ["Simulation Number: %d" % sim_no] if sim_no >= 0 else []
    P.summary(extra)
    init = time.time()
    tag = P.ktag()
    if sim_no != -1:
        pk_file_name = P.outdir + "pk_%s%d_%s_N%d_%s.txt" % (P.string, sim_no, P.weight_type, P.N_mc, tag)
    else:
        pk_file_name = P.outdir + "pk_%s%s_N%d_%s.txt" % (P.string, P.weight_type, P.N_mc, tag)
    if os.path.exists(pk_file_name):
        print("Simulation has already been computed; exiting!")
        sys.exit()
    bias_file_name = lambda r: P.mcdir + "%s%d_%s_pk_q-bar_a_%s.npy" % (P.string, r, P.weight_type, tag)
    fish_file_name = lambda r: P.mcdir + "%s%d_%s_pk_fish_a_%s.npy" % (P.string, r, P.weight_type, tag)
    combined_bias = P.outdir + "bias-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag)
    combined_fish = P.outdir + "fisher-pk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag)
    if not (os.path.exists(combined_bias) and os.path.exists(combined_fish)):
        for i in range(1, P.N_mc + 1):
            if not os.path.exists(bias_file_name(i)):
                raise Exception("Bias term %d not found" % i)
            if not os.path.exists(fish_file_name(i)):
                raise Exception("Fisher matrix %d not found" % i)
    P.unsupported()
    if sim_no != -1:
        print("\n## Analyzing %s simulation %d with %s weights" % (P.string, sim_no, P.weight_type))
    else:
        print("\n## Analyzing %s with %s weights" % (P.string, P.weight_type))
    ctx, mesh, nbar_A, alpha_ran, shot_fac, diff = _setup(P, sim_no, ("pk_data",) if P.wtype == 0 else ("pk_ml",), paint_data=True)
    print("alpha = %.3f, shot_factor: %.3f" % (alpha_ran, shot_fac))
    print("\n## Computing C-inverse of data and associated computations assuming %s weightings\n" % P.weight_type)
    if P.wtype == 0:
        q_alpha = ctx.pk_data_q(diff).cpu().numpy()
    else:   # pk/compute_pk_data.py:194 -- applyCinv(..., rel_tol=1e-6, verb=1, max_it=30)
        P.load_fiducial_pk(ctx)
        q, it = ctx.pk_data_q_ml(diff, max_it=30, rel_tol=1e-6)
        print("Inversion stopped after %d iterations" % it)
        q_alpha = q.cpu().numpy()
    try:
        bias = np.load(combined_bias)
        fish = np.load(combined_fish)
        print("Cleaning up any remaining temporary files")
        for i in range(1, P.N_mc + 1):
            if os.path.exists(fish_file_name(i)):
                os.remove(fish_file_name(i))
            if os.path.exists(bias_file_name(i)):
                os.remove(bias_file_name(i))
    except IOError:
        print("Loading bias term and Fisher matrix from GRF simulations")
        fish, bias = 0.0, 0.0
        for i in range(1, P.N_mc + 1):
            fish += np.load(fish_file_name(i))
            bias += np.load(bias_file_name(i))
        fish /= P.N_mc
        bias /= P.N_mc
        np.save(combined_bias, bias)
        print("Computed bias term from %d realizations and saved to %s" % (P.N_mc, combined_bias))
        np.save(combined_fish, fish)
        print("Computed Fisher matrix from %d realizations and saved to %s\n" % (P.N_mc, combined_fish))
        for i in range(1, P.N_mc + 1):
            os.remove(fish_file_name(i))
            os.remove(bias_file_name(i))
    p_alpha = np.inner(np.linalg.inv(fish), q_alpha - bias)
    write_pk_text(pk_file_name, P, sim_no, p_alpha, mesh.n_k)
    duration = time.time() - init
    print("## Saved output to %s. Exiting after %d seconds (%d minutes)\n" % (pk_file_name, duration, duration // 60))
    sys.exit()


def write_pk_text(path, P, sim_no, p_alpha, n_k):
    """Text layout of pk/compute_pk_data.py:252-281."""
    b, g = P.boxsize_grid, P.grid_3d
    with open(path, "w+") as output:
        if sim_no == -1:
            output.write("####### Power Spectrum of %s #############" % P.string)
        else:
            output.write("####### Power Spectrum of %s Simulation %d #############" % (P.string, sim_no))
        output.write("\n# Weights: %s" % P.weight_type)
        output.write("\n# Fiducial Omega_m: %.3f" % P.OmegaM_fid)
        output.write("\n# Fiducial h: %.3f" % P.h_fid)
        output.write("\n# Forward-model pixellation : %d" % P.include_pix)
        output.write("\n# Random n-bar: %d" % P.rand_nbar)
        output.write("\n# Boxsize: [%.1f, %.1f, %.1f]" % (b[0], b[1], b[2]))
        output.write("\n# Grid: [%d, %d, %d]" % (g[0], g[1], g[2]))
        output.write("\n# k-binning: [%.3f, %.3f, %.3f]" % (P.k_min, P.k_max, P.dk))
        output.write("\n# l-max: %d" % P.lmax)
        output.write("\n# Monte Carlo Simulations: %d" % P.N_mc)
        output.write("\n#")
        if P.lmax == 0:
            output.write("\n# Format: k | P0(k)")
        elif P.lmax == 2:
            output.write("\n# Format: k | P0(k) | P2(k)")
        elif P.lmax == 4:
            output.write("\n# Format: k | P0(k) | P2(k) | P4(k)")
        else:
            output.write("\n# Format: k | P_{multipoles}(k)")
        output.write("\n############################################")
        for i in range(n_k):
            output.write("\n%.4f" % (P.k_min + (i + 0.5) * P.dk))
            for l_i in range(P.lmax // 2 + 1):
                output.write("\t%.8e" % (p_alpha[i + l_i * n_k]))


# ------------------------------------------------------------------------------ bk randoms
BIAS_DTYPE = os.environ.get("SWW_BIAS_DTYPE", "complex64")   # what the reference writes (bk/compute_bk_randoms.py:273)


class GmapWriter:
    """Asynchronous writer of the `*_bk_bias_alpha_*.npz` files (`g_a`, `tilde_g_a`: complex [n_l, n_k, nx, ny, nz],
    bk/compute_bk_randoms.py:273-274).  The maps are cast to the file dtype ON THE DEVICE, copied to one of `depth`
    pinned host buffer pairs on a side stream, and `np.savez`ed by a worker thread, so the Monte-Carlo loop only waits
    when all buffers are still being written (the reference's synchronous np.savez costs seconds per 9 GB seed)."""

    def __init__(self, ctx, depth=2, dtype=None):
        import queue
        import threading
        self.t = engine.torch()
        self.ctx = ctx
        self.dtype = np.dtype(dtype or BIAS_DTYPE)
        self.tdtype = {np.dtype("complex64"): self.t.complex64, np.dtype("complex128"): self.t.complex128}[self.dtype]
        self.side = self.t.cuda.Stream(device=ctx.device)
        self.free = queue.Queue()
        self.todo = queue.Queue()
        self.depth = depth
        self.bufs = None
        self.err = None
        self.th = threading.Thread(target=self._work, daemon=True)
        self.th.start()

    def _work(self):
        while True:
            item = self.todo.get()
            if item is None:
                return
            path, i, ev = item
            try:
                ev.synchronize()
                g, tg = self.bufs[i]
                np.savez(path, g_a=g.numpy(), tilde_g_a=tg.numpy())
            except Exception as ex:      # surfaced by save() / close()
                self.err = ex
            self.free.put(i)

    def save(self, path, g, tg):
        t = self.t
        if self.err:
            raise self.err
        if self.bufs is None:
            self.bufs = [(t.empty(tuple(g.shape), dtype=self.tdtype).pin_memory(), t.empty(tuple(g.shape), dtype=self.tdtype).pin_memory())
                         for _ in range(self.depth)]
            for i in range(self.depth):
                self.free.put(i)
        i = self.free.get()                      # back-pressure: waits for a buffer pair whose file is written
        self.side.wait_stream(self.ctx.stream)
        with t.cuda.stream(self.side):
            gc, tgc = g.to(self.tdtype), tg.to(self.tdtype)
            cast = t.cuda.Event()
            cast.record(self.side)
            self.bufs[i][0].copy_(gc, non_blocking=True)
            self.bufs[i][1].copy_(tgc, non_blocking=True)
            ev = t.cuda.Event()
            ev.record(self.side)
        # the next pair overwrites g / tg on the context stream: it only has to wait for the device-side casts
        self.ctx.stream.wait_event(cast)
        self.todo.put((path, i, ev))

    def close(self):
        self.todo.put(None)
        self.th.join()
        if self.err:
            raise self.err


def save_gmaps(path, g, tg):
    """`g_a`, `tilde_g_a` arrays [n_l, n_k, nx, ny, nz] as in bk/compute_bk_randoms.py:273-274 (synchronous)."""
    t = engine.torch()
    td = t.complex64 if np.dtype(BIAS_DTYPE) == np.dtype("complex64") else t.complex128
    np.savez(path, g_a=g.to(td).cpu().numpy(), tilde_g_a=tg.to(td).cpu().numpy())


def bk_randoms_main(argv):
    """bk/compute_bk_randoms.py: g-maps of a pair of realisations (saved) and their Fisher contribution."""
    rand_it, paramfile = _argv(argv, "Need to specify random iteration and parameter file!")
    P = Params(paramfile, "bk")
    tmpdir = P.tmpdir + str("%d/" % rand_it)
    if not os.path.exists(P.mcdir):
        os.makedirs(P.mcdir)
    if rand_it > P.N_mc // 2:
        raise Exception("Simulation number cannot be greater than half the number of bias sims!")
    rand_it1, rand_it2 = 2 * rand_it, 2 * rand_it + 1
    P.summary(["Random iteration pair: %d, %d" % (rand_it1, rand_it2), "Temporary Directory: %s" % tmpdir])
    init = time.time()
    tag = P.ktag()
    bias_file_name = lambda ii: P.mcdir + "%s%d_%s_bk_bias_alpha_%s.npz" % (P.string, ii, P.weight_type, tag)
    fish_file_name = P.mcdir + "%s%d_%s_bk_fish_alpha_beta_%s.npy" % (P.string, rand_it, P.weight_type, tag)
    if os.path.exists(fish_file_name) and os.path.exists(bias_file_name(rand_it1)) and os.path.exists(bias_file_name(rand_it2)):
        print("Simulation already completed; exiting!\n")
        sys.exit()
    P.unsupported()
    print("\n## Loading %s random pair (%d,%d) with %s weights\n" % (P.string, rand_it1, rand_it2, P.weight_type))
    if os.path.exists(tmpdir):
        shutil.rmtree(tmpdir)      # nothing is spilled on this path; only stale crud is removed
    print("Loading data")
    ctx, mesh, nbar_A, alpha_ran, shot_fac, _ = _setup(P, 1, ("bk_fisher",), paint_data=False)
    print("Data-to-random ratio: %.3f, shot-noise factor: %.3f" % (alpha_ran, shot_fac))
    np.random.seed(rand_it1)
    diffA = np.random.normal(loc=0.0, scale=np.sqrt(nbar_A), size=nbar_A.shape)
    np.random.seed(rand_it2)
    diffB = np.random.normal(loc=0.0, scale=np.sqrt(nbar_A), size=nbar_A.shape)
    print("\n## Computing g-a maps, phi-alpha maps and the Fisher matrix contribution in %d bins on the GPU" % ctx.n_b_bk)
    print("Computing Delta_{abc} normalization")
    F, (gA, tgA, gB, tgB) = ctx.bk_fisher_pair(diffA, diffB)
    wr = GmapWriter(ctx)                      # the two files are written while the other is cast / copied
    wr.save(bias_file_name(rand_it1), gA, tgA)
    wr.save(bias_file_name(rand_it2), gB, tgB)
    np.save(fish_file_name, F.cpu().numpy())
    wr.close()
    duration = time.time() - init
    print("\n### Exiting after %d seconds (%d minutes)\n" % (duration, duration // 60))
    sys.exit()


# ------------------------------------------------------------------------------ bk data
def bk_data_main(argv):
    """bk/compute_bk_data.py: cubic estimator of a catalogue, B_l(k1,k2,k3) text."""
    sim_no, paramfile = _argv(argv, "Need to specify simulation number and parameter file")
    P = Params(paramfile, "bk")
    if not P.use_qbar:
        print("CAUTION: Not including linear term in bispectrum estimator!\n")
    extra = ["Simulation Number: %d" % sim_no] if sim_no >= 0 else []
    P.summary(extra)
    init = time.time()
    tag = P.ktag()
    if sim_no != -1:
        p_alpha_file_name = P.outdir + "bk_%s%d_%s_N%d_%s.txt" % (P.string, sim_no, P.weight_type, P.N_mc, tag)
    else:
        p_alpha_file_name = P.outdir + "bk_%s%s_N%d_%s.txt" % (P.string, P.weight_type, P.N_mc, tag)
    if os.path.exists(p_alpha_file_name):
        print("Simulation has already been computed; exiting!")
        sys.exit()
    P.unsupported()
    if sim_no != -1:
        print("\n## Loading %s simulation %d with %s weights" % (P.string, sim_no, P.weight_type))
    else:
        print("\n## Loading %s with %s weights" % (P.string, P.weight_type))
    ctx, mesh, nbar_A, alpha_ran, shot_fac, diff = _setup(P, sim_no, ("bk_data",), paint_data=True)
    print("alpha = %.3f, shot_factor: %.3f" % (alpha_ran, shot_fac))
    bias_file_name = lambda ii: P.mcdir + "%s%d_%s_bk_bias_alpha_%s.npz" % (P.string, ii, P.weight_type, tag)
    fish_file_name = lambda ii: P.mcdir + "%s%d_%s_bk_fish_alpha_beta_%s.npy" % (P.string, ii, P.weight_type, tag)
    combined_fish = P.outdir + "fisher-bk_%s%d_%s_%s.npy" % (P.string, P.N_mc, P.weight_type, tag)
    try:
        fish = np.load(combined_fish)
        print("Loading Fisher matrix from file")
        print("Cleaning up any remaining temporary files")
        for i in range(1, P.N_mc // 2 + 1):
            if os.path.exists(fish_file_name(i)):
                os.remove(fish_file_name(i))
    except IOError:
        print("Loading Fisher matrix from Gaussian simulations")
        fish = 0.0
        for i in range(1, P.N_mc // 2 + 1):
            fish += np.load(fish_file_name(i))
        fish /= (P.N_mc // 2)
        np.save(combined_fish, fish)
        print("Computed Fisher matrix from %d pairs of realizations and saved to %s\n" % (P.N_mc // 2, combined_fish))
        for i in range(1, P.N_mc // 2 + 1):
            os.remove(fish_file_name(i))
    print("\n## Computing g-a maps assuming %s weightings" % P.weight_type)
    g = ctx.bk_gmaps(diff, data=True)
    print("Computing Delta_{abc} normalization")
    Delta = ctx.bk_delta().cpu().numpy()
    nt, n_l = ctx.n_tri, mesh.n_l
    print("\n## Computing q_alpha quantity in %d bins assuming %s weightings" % (nt * n_l, P.weight_type))
    q = ctx.bk_q3(g)
    if P.use_qbar:
        for ii in range(2, P.N_mc + 2):
            print("1-field: on MC simulation %d of %d" % (ii - 1, P.N_mc))
            mcdat = np.load(bias_file_name(ii))
            g_mc = np.ascontiguousarray(np.real(mcdat["g_a"]), dtype=np.float64)
            tg_mc = np.ascontiguousarray(np.real(mcdat["tilde_g_a"]), dtype=np.float64)
            mcdat.close()
            ctx.bk_q1_accumulate(g, g_mc, tg_mc, P.N_mc, q)
    q = q.cpu().numpy()
    q_alpha = np.concatenate([q[:, l_i] / Delta[l_i * nt:(l_i + 1) * nt] for l_i in range(n_l)])
    print("\n## Computing p_alpha in %d bins assuming %s weightings" % (nt * n_l, P.weight_type))
    p_alpha = np.matmul(np.linalg.inv(fish), q_alpha)
    write_bk_text(p_alpha_file_name, P, sim_no, p_alpha, mesh.triangles())
    duration = time.time() - init
    print("## Saved bispectrum estimates to %s. Exiting after %d seconds (%d minutes)\n" % (p_alpha_file_name, duration, duration // 60))
    sys.exit()


def write_bk_text(path, P, sim_no, p_alpha, bins_index):
    """Text layout of bk/compute_bk_data.py:421-451."""
    b, g = P.boxsize_grid, P.grid_3d
    nt = len(bins_index)
    with open(path, "w+") as output:
        if sim_no != -1:
            output.write("####### Bispectrum of %s Simulation %d #############" % (P.string, sim_no))
        else:
            output.write("####### Bispectrum of %s #############" % (P.string))
        output.write("\n# Weights: %s" % P.weight_type)
        output.write("\n# Fiducial Omega_m: %.3f" % P.OmegaM_fid)
        output.write("\n# Fiducial h: %.3f" % P.h_fid)
        output.write("\n# Boxsize: [%.1f, %.1f, %.1f]" % (b[0], b[1], b[2]))
        output.write("\n# Grid: [%d, %d, %d]" % (g[0], g[1], g[2]))
        output.write("\n# k-binning: [%.3f, %.3f, %.3f]" % (P.k_min, P.k_max, P.dk))
        output.write("\n# l-max: %d" % P.lmax)
        output.write("\n# Monte Carlo Simulations: %d" % P.N_mc)
        output.write("\n#")
        if P.lmax == 0:
            output.write("\n# Format: k1 | k2 | k3 | B0(k1,k2,k3)")
        elif P.lmax == 2:
            output.write("\n# Format: k1 | k2 | k3 | B0(k1,k2,k3) | B2(k1,k2,k3)")
        elif P.lmax == 4:
            output.write("\n# Format: k1 | k2 | k3 | B(k1,k2,k3) | B2(k1,k2,k3) | B4(k1,k2,k3)")
        else:
            output.write("\n# Format: k1 | k2 | k3 | B_{multipoles}(k1,k2,k3)")
        output.write("\n############################################")
        k_av = np.arange(P.k_min, P.k_max, P.dk) + P.dk / 2.0
        for i in range(nt):
            a, bb, c = bins_index[i]
            output.write("\n%.4f\t%.4f\t%.4f" % (k_av[a], k_av[bb], k_av[c]))
            for l_i in range(P.lmax // 2 + 1):
                output.write("\t%.8e" % (p_alpha[l_i * nt + i]))
