"""
GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI,
against (a) the golden chains produced by the unmodified reference and (b) the C oracle
on the same seeded inputs.  Tolerances: chain rows 1e-9 relative (fp64 reassociation +
device/host libm differences); integer quantities (weights, row counts) exact.
"""

import numpy as np
import pytest

from tests.util import flat_from_golden, flat_g6, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9
ATOL = 1e-11


def _engine(fm, n_chains, seed, chain_id0=0, rows_cap=4096, burn_in=0):
    from cobaya_b200.engine import Engine

    return Engine(fm, n_chains=n_chains, seed=seed, chain_id0=chain_id0,
                  rows_cap=rows_cap, burn_in=burn_in)


def test_logpost_matches_reference_known_answers(cuda_lib):
    """cb2_logpost vs Model.logposterior values recorded from the reference (KAT1, KAT5)."""
    u = load_golden("units")
    g = load_golden("g1_gauss3d")
    fm = flat_from_golden(g)
    eng = _engine(fm, 1, 1)
    lp, pr, ll, der = eng.logpost(u["kat1_points"])
    ref = u["kat1_results"]
    for i in range(len(ref)):
        if np.isfinite(ref[i, 0]):
            np.testing.assert_allclose(lp[i], ref[i, 0], rtol=1e-11)
            np.testing.assert_allclose(pr[i], ref[i, 1], rtol=1e-13)
            np.testing.assert_allclose(ll[i, 0], ref[i, 2], rtol=1e-11)
            np.testing.assert_allclose(der[i], ref[i, 3:], rtol=1e-9, atol=1e-12)
        else:
            assert lp[i] == -np.inf and pr[i] == -np.inf
    g2 = load_golden("g2_blocks_mixture")
    fm2 = flat_from_golden(g2)
    eng2 = _engine(fm2, 1, 1)
    lp, pr, ll, der = eng2.logpost(u["kat5_points"])
    ref = u["kat5_results"]
    np.testing.assert_allclose(lp, ref[:, 0], rtol=1e-11)
    np.testing.assert_allclose(pr, ref[:, 1], rtol=1e-12)
    np.testing.assert_allclose(ll[:, 0], ref[:, 2], rtol=1e-11)
    np.testing.assert_allclose(der, ref[:, 3:], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("policy", [0, 4, 1])
@pytest.mark.parametrize("n", [2, 3, 7, 16, 41, 64, 65, 72, 100, 128, 136, 256, 512])
def test_random_so_n_matches_reference_rvs(cuda_lib, n, policy):
    """Device Haar basis vs the reference's numba _rvs on the same normals (golden)."""
    from cobaya_b200.flatmodel import FlatModel

    u = load_golden("units")
    D = n + 1  # block 0 of size 1, block 1 of size n (golden used block=1)
    fm = FlatModel.gaussian(np.zeros(D), np.eye(D) * 0.01, blocks=[[0], list(range(1, D))],
                            oversampling=[1, 1], proposal_cov=np.eye(D) * 0.01)
    eng = _engine(fm, 20, seed=1234)
    # 0: compact-WY DMMA sweep (4 warps for n <= 64, 8 warps + k_normals for n <= 128, N in
    #    global memory for n <= 512),
    # 4: DFMA sweep, 1: general kernel
    eng.set_kernel_policy(policy)
    R = eng.debug_basis(chain=17, block=1, epoch=5)
    if f"son_R_{n}" in u:
        ref = u[f"son_R_{n}"]   # the reference's numba _rvs on the same normals
    else:
        from oracle import oracle as orc
        ref = orc.random_SO_N(n, 1234, 17, 1, 5)
    # rounding grows with the number of reflectors applied
    np.testing.assert_allclose(R, ref, rtol=0, atol=5e-14 * max(1, n // 64))
    np.testing.assert_allclose(R @ R.T, np.eye(n), atol=1e-13 * max(1, n // 64))
    assert abs(np.linalg.det(R) - 1) < 1e-12


CASES = [("g1_gauss3d", 0), ("g1_gauss3d", 7), ("g2_blocks_mixture", 3),
         ("g3_dragging", 1), ("g4_block1d", 5),
         # 72-D, two blocks, two modes: the streamed kernels against the reference's own rows
         ("g7_stream72", 2), ("g7_stream72", 9)]


@pytest.mark.parametrize("name,cid", CASES)
def test_chain_matches_reference_golden(cuda_lib, name, cid):
    """One chain advanced by the CUDA engine reproduces the rows the unmodified
    reference produced with the same Philox draws (tests/golden, oracle/make_golden.py)."""
    g = load_golden(name)
    fm = flat_from_golden(g)
    n = int(g["n_proposals"])
    eng = _engine(fm, 1, seed=int(g["seed"]), chain_id0=cid, burn_in=int(g["burn_in"]))
    eng.set_state(g[f"x0_{cid}"][None, :])
    # deliberately uneven launch sizes: windows must be transparent
    done = 0
    for k in (1, 2, 5, 17, 100, n):
        step = min(k, n - done)
        if step > 0:
            eng.advance(step)
            done += step
    st = eng.get_state()
    if name == "g3_dragging":
        # one parameter of this case is periodic: k_step_drag refuses, the general kernel runs
        # (k_step_drag is checked by the dragging tests against the oracle below)
        assert eng.last_step_kernel() == 0
    ref = g[f"rows_{cid}"]
    rows = eng.rows(0)
    assert st["flags"][0] == 0
    assert rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, 0], ref[:, 0])  # weights are integers
    np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st["x"][0], g[f"final_x_{cid}"], rtol=RTOL, atol=ATOL)
    assert st["weight"][0] == int(g[f"final_weight_{cid}"])


def _oracle_rows(fm, seed, ids, x0, n, burn_in):
    from oracle import oracle as orc

    om = orc.OracleModel(fm)
    out = []
    for i, cid in enumerate(ids):
        ch = orc.OracleChain(om, seed, int(cid), x0[i], burn_in=burn_in)
        rc, rows = ch.advance(n)
        out.append((rc, rows, ch.state()))
    return out


@pytest.mark.parametrize("policy", [0, 1, 2])
@pytest.mark.parametrize("D,n_chains,n", [(8, 37, 300), (64, 16, 200), (21, 9, 150)])
def test_ensemble_matches_oracle(cuda_lib, D, n_chains, n, policy):
    """Many chains, global ids offset (as on rank>0), against the C oracle."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
    rng = np.random.default_rng(5)
    x0 = rng.multivariate_normal(np.zeros(D), cov, size=n_chains)
    id0 = 1000
    eng = _engine(fm, n_chains, seed=77, chain_id0=id0, rows_cap=n)
    eng.set_kernel_policy(policy)
    eng.set_state(x0)
    eng.advance(n)
    # 0 auto -> producer/consumer DMMA kernel, 2 -> single-role DMMA kernel, 1 -> general
    assert eng.last_step_kernel() == {0: 2, 1: 0, 2: 1}[policy]
    st = eng.get_state()
    ref = _oracle_rows(fm, 77, range(id0, id0 + n_chains), x0, n, 0)
    for c in range(n_chains):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]
        assert st["n_accepted"][c] == s_ref["n_accepted"]


@pytest.mark.parametrize("policy,D", [(0, 6), (1, 6), (0, 64), (0, 20), (0, 100), (0, 136)])
def test_moments_match_numpy(cuda_lib, policy, D):
    """cb2_moments (multi-chain halves rule) vs SampleCollection.mean/cov arithmetic
    restated with numpy on the same rows, and R-1 to 1e-4 as BASELINE.json asks."""
    from cobaya_b200.convergence import rminus1_from_sums
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    from oracle import oracle as orc

    C, n = 24, 600
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.full(D, 0.1), cov, proposal_cov=cov)
    rng = np.random.default_rng(2)
    x0 = 0.1 + rng.multivariate_normal(np.zeros(D), cov, size=C)
    eng = _engine(fm, C, seed=5, rows_cap=n)
    eng.set_kernel_policy(policy)
    eng.set_state(x0)
    eng.advance(n)
    shift = np.full(D, 0.09)
    sums = eng.moments(shift=shift)
    Ns, means, covs, accs = [], [], [], []
    for c in range(C):
        rows = eng.rows(c)
        m, cv, a = orc.chain_window_stats(rows, D, len(rows) // 2)
        Ns.append(len(rows)); means.append(m); covs.append(cv); accs.append(a)
    R_ref, W_ref = orc.rminus1_from_chain_stats(Ns, means, covs)
    res = rminus1_from_sums(sums, D, shift)
    assert res["M"] == C and res["N"] == sum(Ns)
    np.testing.assert_allclose(res["W"], W_ref, rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(res["Rminus1"], R_ref, rtol=1e-4)
    np.testing.assert_allclose(res["acceptance"], np.average(accs, weights=Ns), rtol=1e-12)
    np.testing.assert_allclose(res["mean"], np.mean(means, axis=0), rtol=1e-9, atol=1e-12)


def test_single_chain_split_matches_reference_checkpoint(cuda_lib):
    """Single-chain R-1 (mcmc.py:795-822) and learned covariance vs the values the
    reference itself computed on its own chain (tests/golden/checkpoint.npz)."""
    from cobaya_b200.convergence import rminus1_from_sums
    from cobaya_b200.engine import MOMENTS_SINGLE_SPLIT

    ck = load_golden("checkpoint")
    g = load_golden("g1_gauss3d")
    fm = flat_from_golden(g)
    eng = _engine(fm, 1, seed=9, chain_id0=2, rows_cap=4096)
    eng.set_state(ck["x0"][None, :])
    eng.advance(6000)
    rows = eng.rows(0)
    assert rows.shape == ck["rows"].shape
    np.testing.assert_allclose(rows, ck["rows"], rtol=RTOL, atol=ATOL)
    sums = eng.moments(mode=MOMENTS_SINGLE_SPLIT, split=int(ck["split"]))
    res = rminus1_from_sums(sums, 3)
    np.testing.assert_allclose(res["Rminus1"], float(ck["Rminus1"]), rtol=1e-4)
    np.testing.assert_allclose(res["W"], ck["learned_cov"], rtol=1e-9)
    np.testing.assert_allclose(res["acceptance"], float(ck["acceptance"]), rtol=1e-12)


def _mixed_model(D=10, seed=3, modes=2, thin=1):
    """2 blocks with oversampling, 2-mode mixture with weights whose input order differs
    from the block-sorted order (dense likelihood matrix in the DMMA path), one normal
    prior, one periodic parameter."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(seed)
    means, covs = [], []
    for k in range(modes):
        A = rng.standard_normal((D, 2 * D))
        C = A @ A.T / (2 * D)
        d = np.sqrt(np.diag(C))
        s = 0.05 * (1 + rng.uniform(0, 1, D))
        covs.append((C / d[:, None] / d[None, :]) * s[:, None] * s[None, :])
        means.append(rng.uniform(-0.1, 0.1, D) + 0.15 * k)
    idx = rng.permutation(D)
    lk = LikeSpec.gaussian_mixture(idx, np.array(means), np.array(covs),
                                   weights=None if modes == 1 else rng.uniform(0.5, 1, modes),
                                   name="gm")
    kind = np.zeros(D, np.int32); kind[2] = 1
    lower = np.full(D, -1.0); upper = np.full(D, 1.0)
    lower[2], upper[2] = -np.inf, np.inf
    periodic = np.zeros(D, np.int32); periodic[5] = 1
    lower[5], upper[5] = -0.3, 0.3
    loc = np.zeros(D); sc = np.ones(D); sc[2] = 0.4
    blocks = [[7, 0, 3], [1, 2, 4, 5, 6, 8, 9][: D - 3]]
    prop = np.diag(np.full(D, 0.05**2))
    prop[0, 1] = prop[1, 0] = 0.3 * 0.05**2
    return FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=kind, lower=lower,
                     upper=upper, loc=loc, pscale=sc, periodic=periodic, likes=[lk],
                     blocks=blocks, oversampling=[1, 2], proposal_cov=prop, temperature=1.5,
                     output_thin=thin)


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("modes,thin", [(1, 1), (2, 3)])
def test_general_and_dmma_paths_match_oracle(cuda_lib, policy, modes, thin):
    """The same model through the general (warp-per-chain) and the DMMA step kernels,
    both against the oracle; blocks, mixture, normal/periodic priors, thinning, burn-in."""
    fm = _mixed_model(modes=modes, thin=thin)
    C, n = 19, 400
    rng = np.random.default_rng(8)
    x0 = rng.uniform(-0.05, 0.05, (C, fm.D))
    eng = _engine(fm, C, seed=21, chain_id0=300, rows_cap=n, burn_in=3)
    eng.set_kernel_policy(policy)
    eng.set_state(x0)
    for k in (3, 50, 347):
        eng.advance(k)
    # periodic parameter / two modes: the single-role DMMA kernel (1), not the
    # producer/consumer one
    assert eng.last_step_kernel() == (1 if policy == 0 else 0)
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 21, range(300, 300 + C), x0, n, 3)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]


def test_dmma_path_is_used_for_headline_config(cuda_lib):
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    cov = synthetic_gaussian_cov(64)
    fm = FlatModel.gaussian(np.zeros(64), cov, proposal_cov=cov)
    eng = _engine(fm, 8, seed=1, rows_cap=64)
    eng.set_state(np.zeros((8, 64)))
    eng.advance(64)
    assert eng.last_step_kernel() == 2


# ---------------------------------------------------------------------------------------
# The other BASELINE.json configurations as parity cases (engine vs oracle, few chains)
# ---------------------------------------------------------------------------------------
def _mixture_cov(D, rng, scale=0.02):
    A = rng.standard_normal((D, 2 * D))
    C = A @ A.T / (2 * D)
    d = np.sqrt(np.diag(C))
    s = scale * 10 ** rng.uniform(-0.5, 0.5, D)
    return (C / d[:, None] / d[None, :]) * s[:, None] * s[None, :]


def test_config3_128d_three_modes_two_speed_blocks(cuda_lib):
    """configs[2]: 128-D, 3-mode mixture, two components of different speed -> two blocks
    (params 0-31 slow, 32-127 fast, oversampling [1,3], oversample_thin)."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(20260925)
    D, n_slow = 128, 32
    like_a = LikeSpec.gaussian_mixture(
        np.arange(n_slow), [np.full(n_slow, 0.3 * k) * 0.1 for k in range(3)],
        [_mixture_cov(n_slow, rng) for _ in range(3)], name="slow")
    like_b = LikeSpec.gaussian_mixture(
        np.arange(n_slow, D), [np.full(D - n_slow, 0.3 * k) * 0.1 for k in range(3)],
        [_mixture_cov(D - n_slow, rng) for _ in range(3)], name="fast")
    blocks = [list(range(n_slow)), list(range(n_slow, D))]
    cycle = n_slow + 3 * (D - n_slow)
    thin = int(np.round(cycle / D))
    prop = np.diag(np.full(D, 0.02**2))
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -1.0), upper=np.full(D, 1.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[like_a, like_b],
                   blocks=blocks, oversampling=[1, 3], proposal_cov=prop, output_thin=thin)
    assert fm.cycle_length == 320 and thin == 2
    C, n = 6, 700  # more than two proposal cycles
    x0 = rng.normal(0, 0.01, (C, D))
    eng = _engine(fm, C, seed=4, chain_id0=8192 * 7, rows_cap=n)
    eng.set_state(x0)
    eng.advance(333)
    eng.advance(n - 333)
    assert eng.last_step_kernel() == 3  # 64 < D <= 128: streamed kernels, two components
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 4, range(8192 * 7, 8192 * 7 + C), x0, n, 0)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)


def test_config4_30d_rosenbrock_dragging(cuda_lib):
    """configs[3]: 30-D Rosenbrock (builder-defined external-likelihood stand-in), manual
    slow/fast blocking, drag: True."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    D, n_slow, o_fast = 30, 10, 4
    lk = LikeSpec.rosenbrock(np.arange(D), scale=1.0 / 20.0)
    n_drag = int(np.round(o_fast * (D - n_slow) / n_slow))
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -5.0), upper=np.full(D, 5.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=[list(range(n_slow)), list(range(n_slow, D))],
                   oversampling=[1, o_fast], drag=True, i_last_slow_block=0,
                   drag_interp_steps=n_drag, proposal_cov=np.diag(np.full(D, 0.05**2)))
    C, n = 10, 120
    rng = np.random.default_rng(1)
    x0 = 1.0 + rng.normal(0, 0.05, (C, D))
    eng = _engine(fm, C, seed=12, chain_id0=40, rows_cap=n)
    eng.set_state(x0)
    for k in (7, 13, 100):
        eng.advance(k)
    # dragging with the chain state in registers (k_step_drag, Rosenbrock in fragment layout)
    assert eng.last_step_kernel() == 1
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 12, range(40, 40 + C), x0, n, 0)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]


@pytest.mark.parametrize("D,n", [(32, 200), (128, 260), (200, 450), (512, 40), (512, 600)])
def test_config5_dimension_sweep_parity(cuda_lib, D, n):
    """configs[4]: the D sweep as a parity case (single mode, one block)."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
    C = 5
    x0 = np.random.default_rng(D).multivariate_normal(np.zeros(D), cov, size=C)
    eng = _engine(fm, C, seed=D, chain_id0=3, rows_cap=n)
    eng.set_state(x0)
    eng.advance(n)
    # D <= 64: producer/consumer DMMA kernel; above: streamed kernels (DMMA products up to
    # D = 128, cuBLAS GEMMs up to D = 512)
    assert eng.last_step_kernel() == (2 if D <= 64 else 3)
    st = eng.get_state()
    ref = _oracle_rows(fm, D, range(3, 3 + C), x0, n, 0)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)


def test_streamed_kernels_blocks_modes_priors(cuda_lib):
    """The streamed path (k_stream_products / k_stream_whiten / k_stream_accept) for
    64 < D <= 128: three blocks (one of a single parameter) with oversampling and thinning,
    a 3-mode mixture over permuted parameters, a normal and a scipy-family prior, burn-in and
    temperature, several advance calls that cut windows inside proposal cycles."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(11)
    D = 100
    covs = [_mixture_cov(D, rng, scale=0.05) for _ in range(3)]
    means = [rng.uniform(-0.05, 0.05, D) for _ in range(3)]
    lk = LikeSpec.gaussian_mixture(rng.permutation(D), means, covs, weights=[0.5, 0.3, 0.2])
    kind = np.zeros(D, np.int32); kind[3] = 1
    lower = np.full(D, -1.0); upper = np.full(D, 1.0)
    lower[3], upper[3] = -np.inf, np.inf
    sc = np.ones(D); sc[3] = 0.3
    perm = list(rng.permutation(D))
    blocks = [[int(perm[0])], [int(v) for v in perm[1:41]], [int(v) for v in perm[41:]]]
    fm = FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=kind, lower=lower, upper=upper,
                   loc=np.zeros(D), pscale=sc, periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=blocks, oversampling=[1, 2, 3],
                   proposal_cov=np.diag(np.full(D, 0.03**2)), temperature=1.5, output_thin=2)
    C, n = 6, 640
    x0 = rng.uniform(-0.02, 0.02, (C, D))
    eng = _engine(fm, C, seed=17, chain_id0=900, rows_cap=n, burn_in=3)
    eng.set_state(x0)
    for k in (1, 130, 258, 251):
        eng.advance(k)
    assert eng.last_step_kernel() == 3
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 17, range(900, 900 + C), x0, n, 3)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]


def test_streamed_path_takes_two_components_at_small_d(cuda_lib):
    """Two gaussian_mixture components of different speed over disjoint parameters at D = 12
    (the shape of the reference's own speed-blocking tests, common_sampler.py:264-372): the
    register-resident kernels take one component only, so this runs on the streamed kernels."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(23)
    D, ns = 12, 4
    a = LikeSpec.gaussian_mixture(np.arange(ns), [rng.uniform(-0.05, 0.05, ns)],
                                  [_mixture_cov(ns, rng, scale=0.05)], name="slow")
    b = LikeSpec.gaussian_mixture(np.arange(ns, D),
                                  [rng.uniform(-0.05, 0.05, D - ns) for _ in range(2)],
                                  [_mixture_cov(D - ns, rng, scale=0.05) for _ in range(2)],
                                  name="fast")
    fm = FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -1.0), upper=np.full(D, 1.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[a, b],
                   blocks=[list(range(ns)), list(range(ns, D))], oversampling=[1, 3],
                   proposal_cov=np.diag(np.full(D, 0.04**2)), output_thin=2)
    C, n = 9, 400
    x0 = rng.uniform(-0.03, 0.03, (C, D))
    eng = _engine(fm, C, seed=41, chain_id0=5, rows_cap=n)
    eng.set_state(x0)
    for k in (3, 150, 247):
        eng.advance(k)
    assert eng.last_step_kernel() == 3
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 41, range(5, 5 + C), x0, n, 0)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)


def test_producer_consumer_kernel_blocks_normal_prior_thinning(cuda_lib):
    """The producer/consumer DMMA kernel with two blocks + oversampling + thinning, a normal
    prior, burn-in and temperature (one mode, no periodic parameter), against the oracle;
    long enough for the incremental y = L^-1(x-mu) update to be refreshed across windows."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(4)
    D = 12
    cov = _mixture_cov(D, rng, scale=0.05)
    lk = LikeSpec.gaussian_mixture(rng.permutation(D), [rng.uniform(-0.1, 0.1, D)], [cov])
    kind = np.zeros(D, np.int32); kind[3] = 1
    lower = np.full(D, -1.0); upper = np.full(D, 1.0)
    lower[3], upper[3] = -np.inf, np.inf
    sc = np.ones(D); sc[3] = 0.3
    fm = FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=kind, lower=lower, upper=upper,
                   loc=np.zeros(D), pscale=sc, periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=[[5, 1, 9, 0], [2, 3, 4, 6, 7, 8, 10, 11]], oversampling=[1, 2],
                   proposal_cov=np.diag(np.full(D, 0.04**2)), temperature=2.0, output_thin=2)
    C, n = 21, 500
    x0 = rng.uniform(-0.05, 0.05, (C, D))
    eng = _engine(fm, C, seed=31, chain_id0=77, rows_cap=n, burn_in=2)
    eng.set_state(x0)
    for k in (1, 19, 200, 280):
        eng.advance(k)
    assert eng.last_step_kernel() == 2
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 31, range(77, 77 + C), x0, n, 2)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]


@pytest.mark.parametrize("policy", [0, 1])
def test_scipy_priors_match_reference(cuda_lib, policy):
    """All recognised scipy.stats 1-D priors (prior.py:520-525): cb2_logpost against
    Model.logposterior of the reference at 400 points, and one chain against the rows the
    reference produced (g5), on the DMMA step kernel (policy 0) and the general one (1)."""
    g = load_golden("g5_scipy_priors")
    fm = flat_from_golden(g)
    cid, n = 2, int(g["n_proposals"])
    eng = _engine(fm, 1, seed=int(g["seed"]), chain_id0=cid, burn_in=int(g["burn_in"]))
    eng.set_kernel_policy(policy)
    lp, pr, ll, _ = eng.logpost(g["kat_x"])
    ok = np.isfinite(g["kat_logprior"])
    assert 20 < (~ok).sum() < 380
    assert np.all(pr[~ok] == -np.inf) and np.all(lp[~ok] == -np.inf)
    np.testing.assert_allclose(pr[ok], g["kat_logprior"][ok], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ll[ok, 0], g["kat_loglike"][ok], rtol=1e-10)
    eng.set_state(g[f"x0_{cid}"][None, :])
    done = 0
    for k in (3, 64, 500, n):
        step = min(k, n - done)
        eng.advance(step)
        done += step
    assert eng.last_step_kernel() == (1 if policy == 0 else 0)
    st = eng.get_state()
    ref, rows = g[f"rows_{cid}"], eng.rows(0)
    assert st["flags"][0] == 0 and rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, 0], ref[:, 0])
    np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st["x"][0], g[f"final_x_{cid}"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("normalized", [False, True])
def test_gaussian_and_one_likelihoods_match_reference(cuda_lib, normalized):
    """``gaussian`` + ``one`` likelihoods (g6): engine rows vs the reference's."""
    g = load_golden("g6_gaussian_one")
    fm = flat_g6(g, normalized)
    tag = "norm" if normalized else "raw"
    n = int(g["n_proposals"])
    eng = _engine(fm, 1, seed=int(g["seed"]), chain_id0=int(g["chain_id"]))
    eng.set_state(g["x0"][None, :])
    eng.advance(7)
    eng.advance(n - 7)
    ref, rows = g[f"rows_{tag}"], eng.rows(0)
    assert rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, 0], ref[:, 0])
    np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(eng.get_state()["x"][0], g[f"final_x_{tag}"], rtol=RTOL,
                               atol=ATOL)


@pytest.mark.parametrize("case", ["pc", "blocks_general", "dragging"])
def test_snapshot_resume_is_bit_exact(cuda_lib, case):
    """cb2_export_state / cb2_import_state / cb2_load_rows: an engine restored from a
    snapshot continues exactly like the one that was never stopped (state, rows, moments)."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    rng = np.random.default_rng(9)
    if case == "pc":
        D, C = 16, 40
        cov = synthetic_gaussian_cov(D)
        fm = FlatModel.gaussian(np.zeros((1, D)), cov[None], proposal_cov=cov)
        policy, burn = 0, 0
    elif case == "blocks_general":
        g = load_golden("g2_blocks_mixture")
        fm, C, policy, burn = flat_from_golden(g), 9, 1, 3
        D = fm.D
        cov = np.asarray(g["proposal_cov"])
    else:
        g = load_golden("g3_dragging")
        fm, C, policy, burn = flat_from_golden(g), 7, 0, 0
        D = fm.D
        cov = np.asarray(g["proposal_cov"])
    x0 = rng.multivariate_normal(np.zeros(D), cov * 0.01, size=C)
    if case != "pc":
        x0 += np.atleast_2d(g["means"])[0]
    n1, n2 = 333, 278

    def fresh():
        e = _engine(fm, C, seed=77, chain_id0=1000, rows_cap=n1 + n2, burn_in=burn)
        e.set_kernel_policy(policy)
        return e

    a = fresh()
    a.set_state(x0)
    a.advance(n1)
    blob = a.export_state()
    rows = [a.rows(c) for c in range(C)]
    a.advance(n2)
    b = fresh()
    b.import_state(blob, rows)
    b.advance(n2)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    for c in range(C):
        np.testing.assert_array_equal(a.rows(c), b.rows(c))
    np.testing.assert_array_equal(a.moments(), b.moments())
    assert a.last_step_kernel() == b.last_step_kernel()
    # a snapshot of another configuration is refused
    c2 = _engine(fm, C, seed=78, chain_id0=1000, rows_cap=n1 + n2, burn_in=burn)
    with pytest.raises(Exception, match="seed"):
        c2.import_state(blob, rows)


@pytest.mark.parametrize("n,select", [(900, False), (900, True), (90000, False)])
def test_confidence_bounds_match_numpy(cuda_lib, monkeypatch, n, select):
    """cb2_bounds vs the weighted-quantile definition (sort, cumsum, searchsorted) restated
    with numpy on the same rows; Rminus1_cl as mcmc.py:977-982.  Short windows are sorted in
    shared memory (k_task_bounds); windows above 8192 rows -- and every window with
    CB2_BOUNDS_SELECT=1 -- go through the radix selection (k_task_bounds_select): the same
    numbers, no thinning."""
    from cobaya_b200.convergence import rminus1_cl_from_sums, rminus1_from_sums
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    if select:
        monkeypatch.setenv("CB2_BOUNDS_SELECT", "1")
    D, C = 7, 13
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.full(D, 0.1), cov, proposal_cov=cov)
    x0 = 0.1 + np.random.default_rng(3).multivariate_normal(np.zeros(D), cov, size=C)
    eng = _engine(fm, C, seed=8, rows_cap=n)
    eng.set_state(x0)
    eng.advance(n)
    shift = np.full(D, 0.1)
    limfrac = 0.95 / 2
    bs = eng.bounds(limfrac, shift=shift)
    if n > 20000:   # the long case really is longer than the shared-memory sort
        assert eng.get_state()["n_rows"].min() // 2 > 8192
    lows, ups = [], []
    for c in range(C):
        rows = eng.rows(c)
        r = rows[len(rows) // 2:]
        lo, up = [], []
        for i in range(D):
            v, w = r[:, 2 + i], r[:, 0]
            idx = np.argsort(v, kind="stable")
            cs = np.cumsum(w[idx])
            for frac, dst in ((limfrac, lo), (1 - limfrac, up)):
                ix = min(np.searchsorted(cs, cs[-1] * frac), len(v) - 1)
                dst.append(v[idx[ix]])
        lows.append(lo); ups.append(up)
    lows, ups = np.array(lows), np.array(ups)
    assert bs[0] == C
    np.testing.assert_allclose(bs[1:1 + D], (lows - shift).sum(0), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(bs[1 + D:1 + 2 * D], ((lows - shift) ** 2).sum(0), rtol=1e-12)
    np.testing.assert_allclose(bs[1 + 2 * D:1 + 3 * D], (ups - shift).sum(0), rtol=1e-12,
                               atol=1e-15)
    res = rminus1_from_sums(eng.moments(shift=shift), D, shift)
    want = max(np.max(np.std(lows, axis=0) / np.sqrt(np.diag(res["W"]))),
               np.max(np.std(ups, axis=0) / np.sqrt(np.diag(res["W"]))))
    np.testing.assert_allclose(rminus1_cl_from_sums(bs, D, res["W"]), want, rtol=1e-7)


@pytest.mark.parametrize("case", ["pc", "fast_blocks", "streamed", "dragging",
                                  "external_prior", "external_dragging"])
def test_windows_chunked_over_chains_are_bit_identical(cuda_lib, case, monkeypatch):
    """cb2_advance runs a window chunk by chunk over the chains when the per-window buffers
    (Haar bases, streamed products) of all chains do not fit in device memory; the chunks see
    shifted per-chain arrays and chain ids.  Forced here with CB2_CHAIN_CHUNK: the result must
    be identical, bit for bit, to the unchunked run on every kernel family."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    rng = np.random.default_rng(12)
    if case == "pc":
        D, C, n = 64, 72, 300
        cov = synthetic_gaussian_cov(D)
        fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
        x0 = rng.multivariate_normal(np.zeros(D), cov, size=C)
        burn = 0
    elif case == "fast_blocks":
        fm, C, n, burn = _mixed_model(modes=2, thin=3), 40, 333, 3
        x0 = rng.uniform(-0.05, 0.05, (C, fm.D))
    elif case == "streamed":
        D, C, n = 136, 24, 300
        cov = synthetic_gaussian_cov(D)
        fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
        x0 = rng.multivariate_normal(np.zeros(D), cov, size=C)
        burn = 0
    elif case == "external_prior":   # external-function route, prior components per chain
        from tests.test_gpu_external import flat_g9

        fm, C, n, burn = flat_g9(load_golden("g9_external_prior")), 40, 120, 0
        x0 = np.array([0.5, 0.05, 0.1]) + rng.normal(0, 0.02, (C, 3))
    elif case == "external_dragging":   # split-launch dragging with an external Rosenbrock
        from tests import ext_functions

        _, fm, start = ext_functions.rosenbrock_pair()
        C, n, burn = 40, 40, 0
        x0 = start(C, 0)
    else:
        g = load_golden("g3_dragging")
        fm, C, n, burn = flat_from_golden(g), 40, 150, 0
        x0 = np.atleast_2d(g["means"])[0] + rng.normal(0, 0.01, (C, fm.D))

    def run(chunk):
        if chunk:
            monkeypatch.setenv("CB2_CHAIN_CHUNK", str(chunk))
        else:
            monkeypatch.delenv("CB2_CHAIN_CHUNK", raising=False)
        e = _engine(fm, C, seed=3, chain_id0=500, rows_cap=n, burn_in=burn)
        e.set_state(x0)
        e.advance(n // 3)
        e.advance(n - n // 3)
        st = e.get_state()
        rows, counts = e.rows_bulk()
        return st, rows, counts, e.window_counts(), e.last_step_kernel()

    sa, ra, ca, wa, ka = run(0)
    sb, rb, cb, wb, kb = run(16)
    assert wa["chunked_over_chains"] == 0 and wb["chunked_over_chains"] > 0 and ka == kb
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(ra, rb)


@pytest.mark.parametrize("policy", [0, 1])
def test_dragging_mixture_blocks_priors_match_oracle(cuda_lib, policy):
    """get_new_sample_dragging (mcmc.py:564-668) on k_step_drag (policy 0) and on the general
    kernel (policy 1) against the oracle: 2-mode mixture over permuted parameters, two slow and
    two fast blocks (cycler tapes on both sides, a 1-parameter block), a normal and a
    scipy-family prior, a bound that early-rejects some slow proposals, burn-in, thinning off,
    temperature, advance calls that cut windows inside cycles."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    rng = np.random.default_rng(21)
    D = 14
    covs = [_mixture_cov(D, rng, scale=0.05) for _ in range(2)]
    means = [rng.uniform(-0.05, 0.05, D), rng.uniform(-0.05, 0.05, D) + 0.08]
    lk = LikeSpec.gaussian_mixture(rng.permutation(D), means, covs, weights=[0.7, 0.3])
    kind = np.zeros(D, np.int32); kind[2] = 1
    lower = np.full(D, -1.0); upper = np.full(D, 1.0)
    lower[2], upper[2] = -np.inf, np.inf
    lower[0], upper[0] = -0.12, 0.12          # tight: some slow proposals leave the prior
    sc = np.ones(D); sc[2] = 0.3
    blocks = [[0, 5, 9], [3], [1, 2, 4, 6], [7, 8, 10, 11, 12, 13]]
    fm = FlatModel(names=[f"p{i}" for i in range(D)], prior_kind=kind, lower=lower, upper=upper,
                   loc=np.zeros(D), pscale=sc, periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=blocks, oversampling=[1, 1, 3, 3], drag=True, i_last_slow_block=1,
                   drag_interp_steps=5, proposal_cov=np.diag(np.full(D, 0.05 ** 2)),
                   temperature=1.5)
    C, n = 19, 260
    x0 = rng.uniform(-0.04, 0.04, (C, D))
    eng = _engine(fm, C, seed=33, chain_id0=70, rows_cap=n, burn_in=2)
    eng.set_kernel_policy(policy)
    eng.set_state(x0)
    for k in (3, 50, 207):
        eng.advance(k)
    assert eng.last_step_kernel() == (1 if policy == 0 else 0)
    st = eng.get_state()
    assert not st["flags"].any()
    ref = _oracle_rows(fm, 33, range(70, 70 + C), x0, n, 2)
    for c in range(C):
        rc, rows_ref, s_ref = ref[c]
        rows = eng.rows(c)
        assert rows.shape == rows_ref.shape, f"chain {c}"
        np.testing.assert_array_equal(rows[:, 0], rows_ref[:, 0])
        np.testing.assert_allclose(rows, rows_ref, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st["x"][c], s_ref["x"], rtol=RTOL, atol=ATOL)
        assert st["weight"][c] == s_ref["weight"]
