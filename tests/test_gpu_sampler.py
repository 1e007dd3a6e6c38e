"""
GPU tests of the sampler as a whole (run on the B200 box): the standalone driver and the
Cobaya plugin reach the reference's own statistical pass marks on the reference's own test
problems (tests/test_mcmc.py:22-82 KL <= 0.07; tests/test_mcmc.py:85-143 dragging moments
within 0.03), learn the proposal covariance, and converge by R-1.
"""

import numpy as np
import pytest

from tests.util import flat_from_golden, load_golden

pytestmark = pytest.mark.gpu


def _kl(m_true, S_true, m, S):
    Sinv = np.linalg.inv(S)
    d = len(m)
    return 0.5 * (np.trace(Sinv @ S_true) + (m - m_true) @ Sinv @ (m - m_true) - d
                  + np.linalg.slogdet(S)[1] - np.linalg.slogdet(S_true)[1])


def _weighted_mean_cov(rows, D):
    w, X = rows[:, 0], rows[:, 2:2 + D]
    m = np.average(X, axis=0, weights=w)
    return m, np.cov(X.T, fweights=w.astype(np.int64))


@pytest.mark.parametrize("temperature", [1, 2])
def test_reference_3d_problem_reaches_kl_tolerance(cuda_lib, temperature):
    """configs[0]: the 3-D Gaussian of tests/common_sampler.py with the deliberately bad
    initial covmat, burn-in and max_tries of tests/test_mcmc.py:30-47."""
    from cobaya_b200.mcmc import EnsembleMCMC

    g = load_golden("g1_gauss3d")
    fm = flat_from_golden(g)
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)
    fm.set_covariance(np.asarray(g["S0"]) * temperature)
    rng = np.random.default_rng(1)
    C = 64
    x0 = rng.uniform(-1, 1, (C, 3))
    opts = {"seed": 1, "burn_in": "100d", "max_tries": 3000, "temperature": temperature,
            "learn_proposal_Rminus1_max": 30, "Rminus1_stop": 0.005, "chains_per_gpu": C,
            "rows_per_chain": 20000, "max_samples": 15000}
    s = EnsembleMCMC(fm, x0, opts).run()
    assert s.converged
    assert any(c.learned for c in s.progress)
    rows = s.samples(skip_samples=0.5)
    m, S = _weighted_mean_cov(rows, 3)
    S = S / temperature  # samples of p^(1/T): cov scales with T (collection.py:89-143)
    assert _kl(mean, cov, m, S) <= 0.07
    # stored rows reproduce their own chi2 / logpost columns (common_sampler.py:344-372)
    from oracle import oracle as orc

    om = orc.OracleModel(fm)
    for r in rows[rng.integers(0, len(rows), 10)]:
        v, lp, ll, der = om.logpost(r[2:5])
        np.testing.assert_allclose(r[1], -v / temperature, rtol=1e-10)
        np.testing.assert_allclose(r[-1], -2 * ll[0], rtol=1e-10)
        np.testing.assert_allclose(r[5:8], der, rtol=1e-9, atol=1e-12)


def test_dragging_moments(cuda_lib):
    """tests/test_mcmc.py:85-143: 2-D, slow a / fast b, drag: True; mean/std within 0.03."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec
    from cobaya_b200.mcmc import EnsembleMCMC

    # like_a = N(a; 0.2, 0.3... ) in the reference test the two likelihoods are
    # a ~ N(0.2, 0.293) truncated effects included; here: independent Gaussians with the
    # asserted moments, slow block [a], fast block [b]
    la = LikeSpec.gaussian_mixture([0], [[0.2]], [[[0.293**2]]], name="like_a")
    lb = LikeSpec.gaussian_mixture([1], [[0.0]], [[[0.4**2]]], name="like_b")
    fm = FlatModel(names=["a", "b"], prior_kind=[0, 0], lower=[-3, -3], upper=[3, 3],
                   loc=[0, 0], pscale=[1, 1], periodic=[0, 0], likes=[la, lb],
                   blocks=[[0], [1]], oversampling=[1, 6], drag=True, i_last_slow_block=0,
                   drag_interp_steps=6, proposal_cov=np.diag([0.09, 0.16]))
    C = 256
    x0 = np.random.default_rng(3).normal([0.2, 0.0], 0.1, (C, 2))
    s = EnsembleMCMC(fm, x0, {"seed": 1, "chains_per_gpu": C, "Rminus1_stop": 0.002,
                              "rows_per_chain": 4000, "max_samples": 3000}).run()
    rows = s.samples(skip_samples=0.3)
    m, S = _weighted_mean_cov(rows, 2)
    assert abs(m[0] - 0.2) < 0.03 and abs(m[1]) < 0.03
    assert abs(np.sqrt(S[0, 0]) - 0.293) < 0.03 and abs(np.sqrt(S[1, 1]) - 0.4) < 0.03


def test_64d_ensemble_learns_covmat_and_converges(cuda_lib):
    """configs[1] at reduced chain count: R-1 falls below the stop value, the learned
    proposal covariance approaches the truth, ensemble mean/cov match the analytic target
    within Monte-Carlo error."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    from cobaya_b200.mcmc import EnsembleMCMC

    D, C = 64, 1024
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=np.diag(np.diag(cov)))
    x0 = np.random.default_rng(0).multivariate_normal(np.zeros(D), cov, size=C)
    s = EnsembleMCMC(fm, x0, {"seed": 2, "chains_per_gpu": C, "Rminus1_stop": 0.05,
                              "rows_per_chain": 16000}).run()
    assert s.converged and s.Rminus1_last < 0.05
    assert s.engine.last_step_kernel() == 2
    assert any(c.learned for c in s.progress)
    mean, cv, res = s.mean_and_cov()
    sd = np.sqrt(np.diag(cov))
    assert np.max(np.abs(mean) / sd) < 0.1
    assert _kl(np.zeros(D), cov, mean, cv) < 0.07 * D / 3
    learned = s.fm.get_covariance()
    assert np.max(np.abs(np.diag(learned) / np.diag(cov) - 1)) < 0.3


def test_cobaya_plugin_runs_through_cobaya_run(cuda_lib):
    """`cobaya.run.run(info)` with the engine as `sampler: mcmc` (drop-in) on the reference's
    3-D test problem; products are real SampleCollection / DataFrame objects."""
    from tests.refenv import enable_reference

    enable_reference()
    import sys

    from cobaya.run import run

    import cobaya_b200.plugin as plugin

    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)
    info = {
        "likelihood": {"gaussian_mixture": {"means": [mean], "covs": [cov],
                                            "input_params_prefix": "a_",
                                            "output_params_prefix": "", "derived": True}},
        "params": dict({f"a__{i}": {"prior": {"min": -1, "max": 1}} for i in range(3)},
                       **{f"_{i}": None for i in range(3)}),
        "sampler": {"mcmc": {"covmat": np.asarray(g["S0"]), "covmat_params":
                             ["a__0", "a__1", "a__2"], "burn_in": "100d", "max_tries": 3000,
                             "learn_proposal_Rminus1_max": 30, "Rminus1_stop": 0.005,
                             "measure_speeds": False, "seed": 1, "chains_per_gpu": 32,
                             "rows_per_chain": 30000, "max_samples": 20000}},
    }
    saved = sys.modules.get("cobaya.samplers.mcmc")
    try:
        plugin.install_as_mcmc()
        updated, sampler = run(info)
    finally:
        if saved is not None:
            sys.modules["cobaya.samplers.mcmc"] = saved
    assert isinstance(sampler, plugin.MCMC)
    assert sampler.converged
    col = sampler.products(skip_samples=0.5)["sample"]
    m, S = col.mean(), col.cov()
    assert _kl(mean, cov, m, S) <= 0.07
    prog = sampler.products()["progress"]
    assert len(prog) >= 1 and "Rminus1" in prog.columns
    assert len(col) > 1000 and list(col.columns)[:2] == ["weight", "minuslogpost"]


def test_cobaya_run_resume_continues_bit_for_bit(cuda_lib, tmp_path):
    """a run stopped by ``max_samples`` and resumed through ``cobaya.run``
    (``resume=True``) with a larger limit ends exactly where an uninterrupted run with that
    limit ends (engine snapshot next to the chain file; mcmc.py:131-139,187-214)."""
    from tests.refenv import enable_reference

    enable_reference()
    import copy
    import os

    from cobaya.run import run

    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)

    def info(prefix, max_samples):
        return {
            "likelihood": {"gaussian_mixture": {"means": [mean], "covs": [cov],
                                                "input_params_prefix": "a_",
                                                "output_params_prefix": "", "derived": True}},
            "params": dict({f"a__{i}": {"prior": {"min": -1, "max": 1}} for i in range(3)},
                           **{f"_{i}": None for i in range(3)}),
            "sampler": {"cobaya_b200.plugin.MCMC": {
                "covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
                "burn_in": 10, "max_tries": 3000, "learn_proposal_Rminus1_max": 30,
                "Rminus1_stop": 1e-9,
                "measure_speeds": False, "seed": 5, "chains_per_gpu": 16,
                "rows_per_chain": 4000, "max_samples": max_samples}},
            "output": prefix,
        }

    pa, pb = str(tmp_path / "a" / "run"), str(tmp_path / "b" / "run")
    _, full = run(copy.deepcopy(info(pa, 900)), force=True)
    _, first = run(copy.deepcopy(info(pb, 400)), force=True)
    assert os.path.exists(first.snapshot_filename())
    assert 400 <= first.n() < 900
    _, second = run(copy.deepcopy(info(pb, 900)), resume=True)
    assert second.n_steps_raw == full.n_steps_raw
    # same rows, exact values; the ORDER in the collection / chain file is by output segment
    # (one segment per drain: first run, then the resumed part), so compare as sorted sets
    # and chain by chain through the engine
    a, b = full.collection.data.to_numpy(), second.collection.data.to_numpy()
    assert a.shape == b.shape
    key = lambda m: m[np.lexsort(m.T[::-1])]
    np.testing.assert_array_equal(key(a), key(b))
    for c in (0, 7, 15):
        np.testing.assert_array_equal(full._ens.chain_rows(c), second._ens.chain_rows(c))
    assert len(second._segments) == 2 and len(full._segments) == 1
    np.testing.assert_array_equal(full.proposer.get_covariance(),
                                  second.proposer.get_covariance())
    assert len(second.progress) == len(full.progress)
    # the chain file on disk was appended to, not rewritten
    from cobaya.output import load_samples

    back = load_samples(pb, skip=0, combined=True)
    assert len(back) == len(b)
    np.testing.assert_allclose(back.data.to_numpy()[:, :5], b[:, :5], rtol=1e-7)


@pytest.mark.gpu
def test_timed_output_writes_snapshot_and_progress_during_the_run(cuda_lib, tmp_path):
    """mcmc.py:473-481,1045-1078: with ``output_every`` due at every convergence check the
    engine snapshot, ``.progress``, ``.checkpoint`` and ``.covmat`` exist while the run is
    still going (seen from the callback), so a killed run can be resumed."""
    from tests.refenv import enable_reference

    enable_reference()
    import os

    from cobaya.run import run

    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)
    seen = {"snap": 0, "progress_lines": 0}

    def cb(sampler):
        if os.path.exists(sampler.snapshot_filename()):
            seen["snap"] += 1
            with open(sampler.progress_filename()) as f:
                seen["progress_lines"] = max(seen["progress_lines"], len(f.readlines()))
            assert os.path.exists(sampler.checkpoint_filename())

    info = {
        "likelihood": {"gaussian_mixture": {"means": [mean], "covs": [cov],
                                            "input_params_prefix": "a_",
                                            "output_params_prefix": "", "derived": True}},
        "params": dict({f"a__{i}": {"prior": {"min": -1, "max": 1}} for i in range(3)},
                       **{f"_{i}": None for i in range(3)}),
        "sampler": {"cobaya_b200.plugin.MCMC": {
            "covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
            "max_tries": 3000, "learn_proposal_Rminus1_max": 30, "Rminus1_stop": 1e-9,
            "measure_speeds": False, "seed": 5, "chains_per_gpu": 16, "rows_per_chain": 4000,
            "max_samples": 900, "output_every": "0s", "callback_function": cb,
            "callback_every": 1}},
        "output": str(tmp_path / "run"),
    }
    _, smp = run(info, force=True)
    assert seen["snap"] > 0, "no snapshot was written before the run ended"
    assert seen["progress_lines"] >= 2  # header + at least one checkpoint
    with open(smp.progress_filename()) as f:
        lines = f.readlines()
    assert lines[0].startswith("#") and len(lines) == 1 + len(smp.progress)


@pytest.mark.gpu
def test_samples_combined_and_to_getdist_through_the_reference_exporter(cuda_lib, monkeypatch):
    """mcmc.py:1092-1144 / collection.py:1163-1248: ``to_getdist=True`` hands GetDist one
    chain per ensemble chain (skip applied per chain) through the reference's own
    ``SampleCollection.to_getdist``; ``combined=True`` returns all chains in one collection.
    GetDist itself is not installed here: ``MCSamples`` is replaced by a recorder."""
    from tests.refenv import enable_reference

    enable_reference()
    import cobaya.collection as cc
    from cobaya.run import run

    class Recorder:
        def __init__(self, **kw):
            self.kw = kw

    monkeypatch.setattr(cc, "MCSamples", Recorder)
    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)
    info = {
        "likelihood": {"gaussian_mixture": {"means": [mean], "covs": [cov],
                                            "input_params_prefix": "a_",
                                            "output_params_prefix": "", "derived": True}},
        "params": dict({f"a__{i}": {"prior": {"min": -1, "max": 1}} for i in range(3)},
                       **{f"_{i}": None for i in range(3)}),
        "sampler": {"cobaya_b200.plugin.MCMC": {
            "covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
            "max_tries": 3000, "Rminus1_stop": 1e-9, "measure_speeds": False, "seed": 5,
            "chains_per_gpu": 16, "rows_per_chain": 2000, "max_samples": 300}},
    }
    _, smp = run(info)
    mc = smp.samples(to_getdist=True, skip_samples=0.25)
    assert isinstance(mc, Recorder)
    kw = mc.kw
    assert len(kw["samples"]) == 16 and len(kw["weights"]) == 16 and len(kw["loglikes"]) == 16
    ens = smp._ens
    for c in range(16):
        rows = ens.chain_rows(c, 0.25)
        np.testing.assert_array_equal(kw["weights"][c], rows[:, 0])
        np.testing.assert_array_equal(kw["loglikes"][c], rows[:, 1])
        np.testing.assert_array_equal(kw["samples"][c], rows[:, 2:])
    assert kw["names"][:3] == ["a__0", "a__1", "a__2"] and kw["names"][3].endswith("*")
    comb = smp.samples(combined=True, skip_samples=0.25)
    assert len(comb) == sum(len(w) for w in kw["weights"])
    assert smp.products(to_getdist=True, skip_samples=0.25)["sample"].kw["sampler"] == "mcmc"


@pytest.mark.gpu
@pytest.mark.parametrize("D", [72, 160])
def test_streamed_path_learns_and_recovers_the_posterior(cuda_lib, D):
    """The streamed kernels under the full run loop (64 < D <= 128: DMMA products; above:
    cuBLAS products): start from a diagonal proposal, learn the covariance at checkpoints
    (blocked tensor-pipe SYRK), and recover mean and covariance of a correlated Gaussian."""
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    from cobaya_b200.mcmc import EnsembleMCMC

    cov = synthetic_gaussian_cov(D)
    mean = np.linspace(-0.1, 0.1, D)
    fm = FlatModel.gaussian(mean, cov, proposal_cov=np.diag(np.diag(cov)))
    rng = np.random.default_rng(3)
    C = 256
    x0 = mean + rng.multivariate_normal(np.zeros(D), cov, size=C)
    opts = {"seed": 2, "burn_in": 0, "learn_proposal_Rminus1_max": 1e9, "Rminus1_stop": 1e-9,
            "chains_per_gpu": C, "rows_per_chain": 6000, "max_samples": 2500,
            "learn_every": "5d"}
    s = EnsembleMCMC(fm, x0, opts).run()
    assert s.engine.last_step_kernel() == 3
    assert sum(c.learned for c in s.progress) >= 2
    rows = s.samples(skip_samples=0.5)
    m, S = _weighted_mean_cov(rows, D)
    sig = np.sqrt(np.diag(cov))
    # ~3e5 correlated samples: means within a few percent of sigma, variances within 10 %
    assert np.max(np.abs(m - mean) / sig) < 0.08
    assert np.max(np.abs(np.diag(S) / np.diag(cov) - 1)) < 0.12
    # the learned proposal is close to the target: acceptance near the textbook rate
    assert 0.1 < s.progress[-1].acceptance_rate < 0.45
