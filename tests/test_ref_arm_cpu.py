"""The CPU reference arm of bench.py (oracle/ref_arm.py: the reference's low_mem loop in its shipped complex64 precision on
torch-CPU kernels) must compute the same Fisher elements and q-bar as the FP64 oracle, up to single-precision round-off."""
import numpy as np


def test_ref_arm_rows_match_oracle(world):
    from oracle import sww_oracle as orc
    from oracle.ref_arm import PkFisherRefArm
    w = world("toyA")
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
    qbar_o, F_o = orc.pk_fisher_realisation(S, S.grf(1))
    arm = PkFisherRefArm(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, threads=2, n_dots=1)
    arm.prologue()                      # seed 1, like S.grf(1)
    n_k = arm.n_k
    t = arm.t
    picks = [(0, 1), (1, n_k // 2), (arm.n_l - 1, n_k - 1)]
    rows = {al: arm.applyCinv_fkp(arm.applyC_alpha_single(arm.Cinv, *al)) for al in picks}
    maps = {be: arm.applyC_alpha_single(arm.Ainv, *be) for be in picks}
    scale = np.max(np.abs(F_o))
    for (i, a), row in rows.items():
        al = i * n_k + a
        qb = 0.5 * float(t.sum(row * arm.NAinv).real)
        assert abs(qb - qbar_o[al]) < 2e-4 * np.max(np.abs(qbar_o))
        for (j, b), m in maps.items():
            be = j * n_k + b
            f_ab = 0.5 * float(t.sum(row * m).real)
            f_ba = 0.5 * float(t.sum(rows[(j, b)] * maps[(i, a)]).real)
            assert abs(0.5 * (f_ab + f_ba) - F_o[al, be]) < 2e-4 * scale
    # the bounded-sample bookkeeping: two steps = one loop-1 row + one loop-2 row, T_real from the reference's loop counts
    dt1, _, _ = arm.step()
    dt2, T, parts = arm.step()
    assert parts["n_row1"] == 1 and parts["n_row2"] == 1
    expect = parts["t_prologue"] + arm.n_b * (parts["t_row1"] + parts["t_row2"]) + arm.n_b ** 2 * parts["t_dot"]
    assert abs(T - expect) < 1e-9 * expect and T > dt1 + dt2
    arm.close()
