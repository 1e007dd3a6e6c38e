"""
GPU parity at the LAUNCH SHAPES the benchmark numbers are quoted on (run on the B200 box).

The step kernels pick their CTA shape from the number of chains
(``launch_step_pc_t``: wpc = ceil(tiles / SMs) producer/consumer warp pairs per CTA, ring
slices at ``pair * 3 * SLOT``; ``launch_step_fast_t`` likewise; the streamed accept kernel
one warp per chain over several waves).  The small-ensemble parity tests all run with one
pair per CTA, so these tests run the headline shapes themselves -- 2048, 8192 and 65 536
chains at D = 64 (wpc = 2, 7 and 7 with several waves) -- for two windows of 256
proposals and compare a strided sample of chains (every pair index of several CTAs, the
first and the last tile) with the C oracle: rows to 1e-9, integer weights and counters
exact (mcmc.py:545-562,685-748).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9
ATOL = 1e-11
SMS = 148


def _engine(fm, n_chains, seed, chain_id0=0, rows_cap=4096, burn_in=0):
    from cobaya_b200.engine import Engine

    return Engine(fm, n_chains=n_chains, seed=seed, chain_id0=chain_id0,
                  rows_cap=rows_cap, burn_in=burn_in)


def _sample_chains(n_chains, wpc, extra=()):
    """Local chain indices covering every pair index 0..wpc-1 of the first, a middle and
    the last CTA (8 chains per pair: different quads), plus the very first/last chain."""
    tiles = (n_chains + 7) // 8
    ctas = (tiles + wpc - 1) // wpc
    pick = {0, n_chains - 1}
    for cta in sorted({0, 1, ctas // 2, ctas - 2, ctas - 1}):
        if cta < 0 or cta >= ctas:
            continue
        for pair in range(wpc):
            tile = cta * wpc + pair
            for q in ((3 * pair + cta) % 8, (5 * pair + 1) % 8):
                c = tile * 8 + q
                if c < n_chains:
                    pick.add(c)
    pick.update(int(c) for c in extra if 0 <= c < n_chains)
    return sorted(pick)


def _compare_with_oracle(eng, fm, seed, id0, x0, chains, n, burn_in=0, state=None):
    from oracle import oracle as orc

    om = orc.OracleModel(fm)
    st = state or eng.get_state()
    assert not st["flags"].any(), "engine flagged chains"
    worst = 0.0
    for c in chains:
        ch = orc.OracleChain(om, seed, id0 + c, x0[c], burn_in=burn_in)
        rc, ref = ch.advance(n)
        s_ref = ch.state()
        rows = eng.rows(c, n=n)
        assert rows.shape == ref.shape, f"chain {c}: {rows.shape} vs {ref.shape}"
        np.testing.assert_array_equal(rows[:, 0], ref[:, 0], err_msg=f"weights, chain {c}")
        np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL, err_msg=f"chain {c}")
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"], f"chain {c}"
        assert st["n_accepted"][c] == s_ref["n_accepted"], f"chain {c}"
        if len(ref):
            worst = max(worst, float(np.abs(rows - ref).max()))
    return worst


@pytest.mark.parametrize("variant", ["split_products", "producer_per_tile"])
@pytest.mark.parametrize("n_chains", [2048, 8192, 65536])
def test_headline_shape_pc_kernel_matches_oracle(cuda_lib, n_chains, variant):
    """BASELINE configs[1] exactly as bench.py runs it (64-D, one block, proposal = diag of
    the target, bounds [-1, 1]) on both producer/consumer kernels: k_step_pc2 (products
    split over the SM sub-partitions, 2 and 7 tiles per CTA) and k_step_pc (one warp pair
    per tile, 2 and 7 pairs per CTA; policy +8); at 65 536 chains several waves of CTAs; two
    windows of 4 proposal cycles."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    D, n = 64, 512
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, bounds=(-1.0, 1.0),
                            proposal_cov=np.diag(np.diag(cov)))
    rng = np.random.default_rng([1, n_chains])
    x0 = rng.multivariate_normal(np.zeros(D), cov, size=n_chains)
    id0 = 8192 * 3  # as on rank 3 of a multi-GPU run
    eng = _engine(fm, n_chains, seed=1, chain_id0=id0, rows_cap=n)
    if variant == "producer_per_tile":
        eng.set_kernel_policy(8)
    eng.set_state(x0)
    eng.advance(256)
    eng.advance(256)
    assert eng.last_step_kernel() == 2, eng.debug_message()
    assert eng.debug_message() == ""
    wc = eng.window_counts()
    assert wc["dmma-producer-consumer"] == 2 and wc["pc_launch_refused"] == 0
    assert wc["of_pc_split_products"] == (2 if variant == "split_products" else 0)
    tiles = (n_chains + 7) // 8
    wpc = min(7, -(-tiles // SMS))
    assert wpc == {2048: 2, 8192: 7, 65536: 7}[n_chains]
    chains = _sample_chains(n_chains, wpc)
    assert len(chains) >= 20
    _compare_with_oracle(eng, fm, 1, id0, x0, chains, n)


def _two_mode_model(D, rng):
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    def _cov(scale):
        A = rng.standard_normal((D, 2 * D))
        C = A @ A.T / (2 * D)
        d = np.sqrt(np.diag(C))
        s = scale * 10 ** rng.uniform(-0.3, 0.3, D)
        return (C / d[:, None] / d[None, :]) * s[:, None] * s[None, :]

    covs = [_cov(0.04), _cov(0.05)]
    means = [rng.uniform(-0.05, 0.05, D), rng.uniform(-0.05, 0.05, D) + 0.1]
    lk = LikeSpec.gaussian_mixture(rng.permutation(D), means, covs, weights=[0.6, 0.4])
    kind = np.zeros(D, np.int32); kind[1] = 1
    lower = np.full(D, -1.0); upper = np.full(D, 1.0)
    lower[1], upper[1] = -np.inf, np.inf
    sc = np.ones(D); sc[1] = 0.5
    periodic = np.zeros(D, np.int32); periodic[4] = 1
    lower[4], upper[4] = -0.4, 0.4
    half = D // 2
    return FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=kind, lower=lower,
                     upper=upper, loc=np.zeros(D), pscale=sc, periodic=periodic, likes=[lk],
                     blocks=[list(range(half)), list(range(half, D))], oversampling=[1, 2],
                     proposal_cov=np.diag(np.full(D, 0.03 ** 2)), output_thin=1)


def test_headline_shape_single_role_kernel_matches_oracle(cuda_lib):
    """k_step_fast with several warps per CTA (4096 chains -> 512 tiles -> 4 warps per CTA):
    2-mode mixture over permuted parameters, two blocks with oversampling, one normal prior
    and one periodic parameter."""
    rng = np.random.default_rng(77)
    D, C, n = 32, 4096, 300
    fm = _two_mode_model(D, rng)
    x0 = rng.uniform(-0.03, 0.03, (C, D))
    eng = _engine(fm, C, seed=9, chain_id0=500, rows_cap=n, burn_in=2)
    eng.set_state(x0)
    for k in (7, 150, 143):
        eng.advance(k)
    assert eng.last_step_kernel() == 1
    wpc = min(8, -(-(C // 8) // SMS))
    assert wpc == 4
    chains = _sample_chains(C, wpc)
    _compare_with_oracle(eng, fm, 9, 500, x0, chains, n, burn_in=2)


@pytest.mark.parametrize("D", [72, 128])
def test_headline_shape_streamed_kernels_match_oracle(cuda_lib, D):
    """k_stream_products1 / k_stream_whiten / k_stream_accept at 8192 chains (the shape of
    the D > 64 cells of the dimension sweep), one block, single mode."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    C, n = 8192, 2 * D + 9
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, bounds=(-1.0, 1.0), proposal_cov=cov)
    rng = np.random.default_rng(D)
    x0 = rng.multivariate_normal(np.zeros(D), cov, size=C)
    eng = _engine(fm, C, seed=D, chain_id0=11, rows_cap=n)
    eng.set_state(x0)
    eng.advance(D + 5)
    eng.advance(n - D - 5)
    assert eng.last_step_kernel() == 3
    chains = sorted(set(range(0, C, 257)) | {C - 1, C - 2, 4095, 4096})
    _compare_with_oracle(eng, fm, D, 11, x0, chains, n)


def test_bulk_rows_equal_per_chain_rows_at_headline_shape(cuda_lib):
    """cb2_copy_rows_bulk (device-side compaction + one transfer) returns exactly the rows
    cb2_copy_rows returns chain by chain, for row ranges that differ per chain."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    D, C, n = 64, 8192, 256
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, bounds=(-1.0, 1.0), proposal_cov=cov)
    x0 = np.random.default_rng(3).multivariate_normal(np.zeros(D), cov, size=C)
    eng = _engine(fm, C, seed=2, rows_cap=n)
    eng.set_state(x0)
    eng.advance(128)
    first = eng.get_state()["n_rows"].copy()
    rows_a, counts_a = eng.rows_bulk()           # everything so far
    assert counts_a.sum() == first.sum() and np.array_equal(counts_a, first)
    eng.advance(128)
    rows_b, counts_b = eng.rows_bulk(first=first)  # only the rows added since
    total = eng.get_state()["n_rows"]
    assert np.array_equal(counts_b, total - first)
    offs_a = np.concatenate([[0], np.cumsum(counts_a)])
    offs_b = np.concatenate([[0], np.cumsum(counts_b)])
    for c in list(range(0, C, 331)) + [C - 1]:
        full = eng.rows(c)
        got = np.concatenate([rows_a[offs_a[c]:offs_a[c + 1]], rows_b[offs_b[c]:offs_b[c + 1]]])
        np.testing.assert_array_equal(got, full)
