"""Several contexts on several streams (realisations kept in flight side by side, as bench.py and mc_driver.py do) must
not share mutable state: results are bitwise those of one context working alone."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_contexts_in_flight_match_serial(world):
    import torch as t
    from spectra_without_windows_b200 import engine
    w = world("toyD")
    dev = t.device("cuda", 0)
    pk_mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    bk_mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.bk_bins)
    base_pk = engine.Context(pk_mesh, 0, ("pk_fisher",))
    nbar_A = base_pk.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    base_bk = engine.Context(bk_mesh, 0, ("bk_fisher",))
    base_bk.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    rng = np.random.default_rng(11)
    fields = [base_pk.dev(rng.normal(size=tuple(w.grid)) * np.sqrt(nbar_A)) for _ in range(4)]
    # serial references
    ref_pk = [tuple(x.clone() for x in base_pk.pk_fisher(a)) for a in fields]
    ref_bk = [base_bk.bk_fisher_pair(fields[i], fields[i + 1])[0].clone() for i in (0, 2)]
    t.cuda.synchronize()
    streams = [t.cuda.Stream(device=dev) for _ in range(2)]
    pk = [base_pk.clone_for_stream(st, ("pk_fisher",)) for st in streams]
    bk = [base_bk.clone_for_stream(st, ("bk_fisher",)) for st in streams]
    t.cuda.synchronize()
    for rep in range(3):
        got_pk, got_bk = [None] * 4, [None] * 2
        for i in range(4):
            k = i % 2
            with t.cuda.stream(streams[k]):
                got_pk[i] = pk[k].pk_fisher(fields[i])
                if i % 2 == 0:        # a bispectrum pair (DMMA Gram with split-K scratch) right behind it, other stream busy
                    got_bk[i // 2] = bk[(k + 1) % 2].bk_fisher_pair(fields[i], fields[i + 1])[0]
        t.cuda.synchronize()
        for i in range(4):
            assert t.equal(got_pk[i][0], ref_pk[i][0]) and t.equal(got_pk[i][1], ref_pk[i][1]), (rep, i)
        for j in range(2):
            assert t.equal(got_bk[j], ref_bk[j]), (rep, j)
    for c in pk + bk + [base_pk, base_bk]:
        c.close()
