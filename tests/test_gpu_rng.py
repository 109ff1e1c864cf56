"""Device-side numpy-legacy Gaussian stream vs numpy itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_legacy_gaussian_stream_matches_numpy(world):
    from spectra_without_windows_b200 import engine
    w = world("toyC")
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = engine.get_context(mesh, 0, ("ops", "pk_fisher", "pk_data"))
    rng = np.random.default_rng(0)
    for seed, shape in [(1, (20, 18, 15)), (2, (7,)), (12345, (64, 64, 128)), (4294967295, (33, 1, 5)), (77, (200, 160, 170))]:
        sigma = rng.random(shape) + 0.25
        np.random.seed(seed)
        ref = np.random.normal(loc=0.0, scale=sigma, size=sigma.shape)
        got = ctx.grf_legacy(seed, sigma).cpu().numpy()
        assert got.shape == ref.shape
        # identical stream positions (same accept/reject decisions): values agree to the last bits
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-14, seed
        assert np.mean(got == ref) > 0.5


def test_fisher_from_device_field_matches_host_field(world):
    """A Fisher matrix from the device-generated field agrees with the host-RNG one far below 1e-9."""
    from spectra_without_windows_b200 import engine
    w = world("toyD")
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = engine.Context(mesh, 0, ("pk_fisher",))
    nbar_A = ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    sig = np.sqrt(nbar_A)
    np.random.seed(3)
    a_host = np.random.normal(loc=0.0, scale=sig, size=sig.shape)
    F0, q0 = ctx.pk_fisher(a_host)
    F1, q1 = ctx.pk_fisher(ctx.grf_legacy(3, sig))
    e = float((F1 - F0).abs().max() / F0.abs().max())
    assert e < 1e-12
    ctx.close()


def test_batched_fields_and_pipeline_match_numpy(world):
    """Up to 8 independent legacy streams advance side by side; the overlapped pipeline hands out the same fields."""
    from spectra_without_windows_b200 import engine
    w = world("toyD")
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = engine.Context(mesh, 0, ("pk_fisher",))
    sigma = np.random.default_rng(1).random(tuple(w.grid)) + 0.5
    seeds = [5, 6, 7, 4000000000, 9, 10, 11]

    def ref(s):
        np.random.seed(s)
        return np.random.normal(loc=0.0, scale=sigma, size=sigma.shape)

    outs = ctx.grf_legacy_batch(seeds, sigma)
    for s, o in zip(seeds, outs):
        assert np.max(np.abs(o.cpu().numpy() - ref(s)) / np.maximum(np.abs(ref(s)), 1e-300)) < 1e-14, s
    with pytest.raises(Exception, match="batch"):
        ctx.grf_legacy_batch(list(range(9)), sigma)
    for batch in (1, 3):
        got = {}
        for s, a in ctx.field_pipeline(seeds, sigma, batch=batch):
            got[s] = a.clone()          # the buffer is recycled two batches later
        assert sorted(got) == sorted(seeds)
        for s in seeds:
            assert np.max(np.abs(got[s].cpu().numpy() - ref(s)) / np.maximum(np.abs(ref(s)), 1e-300)) < 1e-14, (batch, s)
    ctx.close()


def test_jump_ahead_substreams_are_bit_identical_to_the_sequential_walker(world):
    """One legacy stream split over 128 CTAs per chunk (MT19937 jump-ahead, csrc/rng.cu) must emit exactly the words the
    single-CTA walker emits: same field bit for bit, across chunk boundaries and for a batch of seeds."""
    import os
    from spectra_without_windows_b200 import engine
    w = world("toyD")
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = engine.Context(mesh, 0, ("ops",))
    sigma = np.random.default_rng(4).random((190, 170, 180)) + 0.5      # 5.8e6 values: three chunks of 2.56e6 attempts
    seeds = [3, 4000000001, 77]
    fast = [o.cpu().numpy() for o in ctx.grf_legacy_batch(seeds, sigma)]
    os.environ["SWW_RNG_JUMP"] = "0"
    try:
        slow = [o.cpu().numpy() for o in ctx.grf_legacy_batch(seeds, sigma)]
    finally:
        del os.environ["SWW_RNG_JUMP"]
    for a, b in zip(fast, slow):
        assert np.array_equal(a, b)
    np.random.seed(77)
    ref = np.random.normal(loc=0.0, scale=sigma, size=sigma.shape)
    assert np.max(np.abs(fast[2] - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-14
    ctx.close()
