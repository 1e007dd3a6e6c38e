"""
External likelihood functions through the device-functor route
(cb2_add_external_likelihood, csrc/kernels_ext.cuh, csrc/ext_functor.inl) against the
UNMODIFIED reference running the Python callable (LikelihoodExternalFunction,
cobaya/likelihood.py:150-255): tests/golden/g8_external.npz, made by oracle/make_golden.py
from tests/ext_functions.py.  Floating point: rows to 1e-9 (the CUDA and the numpy versions of
the function differ in the last bits), integer weights exact.
"""

import numpy as np
import pytest

from tests.util import flat_g8, load_golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-11


def _engine(fm, n, seed, id0=0, rows_cap=4096):
    from cobaya_b200.engine import Engine

    return Engine(fm, n_chains=n, seed=seed, chain_id0=id0, rows_cap=rows_cap)


def test_external_logposterior_matches_reference_known_answers(cuda_lib):
    g = load_golden("g8_external")
    eng = _engine(flat_g8(g), 1, 1)
    lp, pr, ll, _ = eng.logpost(g["kat_x"])
    kat = g["kat"]
    finite = np.isfinite(kat[:, 0])
    assert finite.sum() >= 30 and (~finite).sum() >= 1   # some points outside the prior box
    np.testing.assert_array_equal(np.isfinite(lp), finite)
    np.testing.assert_allclose(lp[finite], kat[finite, 0], rtol=1e-12)
    np.testing.assert_allclose(pr[finite], kat[finite, 1], rtol=1e-12)
    np.testing.assert_allclose(ll[finite], kat[finite, 2:], rtol=1e-12, atol=1e-13)


@pytest.fixture(params=["split", "fused"])
def route(request, monkeypatch):
    """The two device routes of external likelihood functions: split launches around the user
    kernels (CB2_EXT_FUSED=0) and the general step kernel recompiled with the functions inlined
    (CB2_EXT_FUSED=1; the default for runs of >= 1024 chains)."""
    monkeypatch.setenv("CB2_EXT_FUSED", "1" if request.param == "fused" else "0")
    return request.param


def _check_route(eng, route):
    rc = eng.ext_route_counts()
    if route == "fused":
        assert rc["fused_loaded"] and rc["fused_windows"] > 0 and not rc["fused_failed"], rc
    else:
        assert rc["fused_windows"] == 0 and not rc["fused_loaded"], rc


@pytest.mark.parametrize("cid", [0, 6])
def test_external_chain_matches_reference_golden(cuda_lib, cid, route):
    g = load_golden("g8_external")
    fm = flat_g8(g)
    n = int(g["n_proposals"])
    eng = _engine(fm, 1, seed=int(g["seed"]), id0=cid)
    eng.set_state(g[f"x0_{cid}"][None, :])
    done = 0
    for k in (1, 3, 40, n):   # windows are transparent
        step = min(k, n - done)
        if step > 0:
            eng.advance(step)
            done += step
    st = eng.get_state()
    assert st["flags"][0] == 0
    ref, rows = g[f"rows_{cid}"], eng.rows(0)
    assert rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, 0], ref[:, 0])
    np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st["x"][0], g[f"final_x_{cid}"], rtol=RTOL, atol=ATOL)
    assert st["weight"][0] == int(g[f"final_weight_{cid}"])
    _check_route(eng, route)


def test_external_route_equals_builtin_kernels_on_the_same_function(cuda_lib, route):
    """A Gaussian written as an external CUDA function walks like the built-in Gaussian on the
    general kernel (same Philox draws; the log-likelihoods differ in summation order only)."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec, synthetic_gaussian_cov

    from tests import ext_functions

    D, C, n = 8, 96, 400
    builtin, ext, mu, cov = ext_functions.gaussian_pair(D)
    x0 = np.random.default_rng(4).multivariate_normal(mu, cov, size=C)
    a, b = _engine(builtin, C, 21), _engine(ext, C, 21)
    a.set_kernel_policy(1)   # general kernel
    for e in (a, b):
        e.set_state(x0)
        e.advance(n)
    sa, sb = a.get_state(), b.get_state()
    assert not sb["flags"].any()
    np.testing.assert_array_equal(sa["n_rows"], sb["n_rows"])
    np.testing.assert_array_equal(sa["weight"], sb["weight"])
    np.testing.assert_allclose(sa["x"], sb["x"], rtol=1e-10, atol=1e-13)
    ra, ca = a.rows_bulk()
    rb, cb = b.rows_bulk()
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(ra[:, 0], rb[:, 0])
    np.testing.assert_allclose(ra, rb, rtol=1e-10, atol=1e-12)
    # the checkpoint statistics see the same chains
    np.testing.assert_allclose(a.moments(), b.moments(), rtol=1e-9, atol=1e-14)
    _check_route(b, route)


def test_external_errors_are_loud(cuda_lib, route):
    from cobaya_b200.engine import EngineError
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    def model(src, name="f", drag=False):
        kw = dict(blocks=[[0], [1]], oversampling=[1, 2], drag=True, i_last_slow_block=0,
                  drag_interp_steps=2) if drag else {}
        return FlatModel(names=["a", "b"], prior_kind=np.zeros(2, np.int32),
                         lower=np.full(2, -1.0), upper=np.full(2, 1.0), loc=np.zeros(2),
                         pscale=np.ones(2), periodic=np.zeros(2, np.int32),
                         likes=[LikeSpec.external([0, 1], src, name)],
                         proposal_cov=np.eye(2) * 0.01, **kw)

    ok = 'extern "C" __device__ double f(const double *p, int n) { return -p[0]*p[0]-p[1]*p[1]; }'
    with pytest.raises(EngineError, match="undefined|error"):
        _engine(model('extern "C" __device__ double f(const double *p, int n) { return zz; }'),
                4, 1)
    e = _engine(model(ok, drag=True), 4, 1)   # dragging with an external function: supported
    e.set_state(np.zeros((4, 2)))
    e.advance(5)
    assert not e.get_state()["flags"].any()
    # NaN from the function: the chain is flagged (the reference raises)
    nan = ('extern "C" __device__ double f(const double *p, int n) '
           '{ return p[0] > 0.05 ? nan("") : -50.0 * (p[0]*p[0] + p[1]*p[1]); }')
    e = _engine(model(nan), 64, 3)
    e.set_state(np.zeros((64, 2)))
    e.advance(200)
    from cobaya_b200.engine import FLAG_INTERNAL

    assert (e.get_state()["flags"] & FLAG_INTERNAL).any()


def test_external_function_through_cobaya_run(cuda_lib, tmp_path):
    """The reference's own input with ``external: <callable>``: the plugin lowers the callable's
    CUDA twin, checks it against the Python function and samples."""
    import copy

    from tests.refenv import enable_reference

    enable_reference()
    from cobaya.log import LoggedError
    from cobaya.run import run

    from tests import ext_functions
    from cobaya_b200.functor import device_function

    info, _ = ext_functions.info_g8()
    opts = dict(info["sampler"]["mcmc"])
    opts.update(chains_per_gpu=256, max_samples=300, learn_proposal=True, seed=3,
                Rminus1_stop=0.0)
    info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    _, smp = run(copy.deepcopy(info))
    rows = smp.products()["sample"]
    assert len(rows) >= 256 * 300
    x = rows[["a", "b", "c"]].to_numpy()
    w = rows["weight"].to_numpy()
    mean = (w[:, None] * x).sum(0) / w.sum()
    # c ~ N(0.1, 0.2) x N(0, 1) prior -> mean 0.1 / 1.04; b follows 2 a^2
    assert abs(mean[2] - 0.1 / 1.04) < 0.02
    assert abs(mean[1] - 2 * (w * x[:, 0] ** 2).sum() / w.sum()) < 0.02
    assert np.allclose(rows["chi2__banana"].to_numpy(),
                       -2 * ext_functions.banana(x[:, 0], x[:, 1]), rtol=1e-9, atol=1e-9)
    # a CUDA twin that computes something else is refused before sampling
    bad = device_function(ext_functions.BANANA_CUDA.replace("0.09", "0.08"))(
        lambda a, b: ext_functions.banana(a, b))
    info2 = copy.deepcopy(info)
    info2["likelihood"]["banana"] = {"external": bad}
    with pytest.raises(LoggedError, match="disagrees"):
        run(info2)


def test_speeds_are_measured_on_the_device(cuda_lib):
    """cb2_measure_speeds (Model.measure_and_set_speeds, model.py:1543-1592): every component
    alone, in evaluations per second; a heavier function is slower; the plugin feeds the
    measured speeds to the reference's automatic blocking."""
    import copy

    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    heavy = ('extern "C" __device__ double heavy(const double *p, int n) { double s = 0.0; '
             'for (int i = 0; i < 20000; ++i) s += sin(p[0] + i * 1e-3) * 1e-9; '
             'return -p[0] * p[0] - p[1] * p[1] + s; }')
    light = ('extern "C" __device__ double light(const double *p, int n) '
             '{ return -p[0] * p[0]; }')
    fm = FlatModel(names=["a", "b", "c"], prior_kind=np.zeros(3, np.int32),
                   lower=np.full(3, -1.0), upper=np.full(3, 1.0), loc=np.zeros(3),
                   pscale=np.ones(3), periodic=np.zeros(3, np.int32),
                   likes=[LikeSpec.external([0, 1], heavy, "heavy", name="heavy"),
                          LikeSpec.external([2], light, "light", name="light"),
                          LikeSpec.gaussian([2], [0.0], [[0.04]], name="gaussian")],
                   proposal_cov=np.eye(3) * 0.01)
    eng = _engine(fm, 1, 1)
    X = np.random.default_rng(0).uniform(-0.5, 0.5, (4096, 3))
    sp = eng.measure_speeds(X, repeats=3)
    assert sp.shape == (3,) and np.all(np.isfinite(sp)) and np.all(sp > 0)
    assert sp[0] < sp[1] / 20 and sp[0] < sp[2] / 20
    # the measurement leaves the posterior untouched
    lp, pr, ll, _ = eng.logpost(X[:4])
    assert np.allclose(ll[:, 1], -X[:4, 2] ** 2)

    from tests.refenv import enable_reference

    enable_reference()
    from cobaya.run import run

    from tests import ext_functions

    info, _ = ext_functions.info_g8()
    opts = dict(info["sampler"]["mcmc"])
    opts.pop("blocking")
    opts.update(chains_per_gpu=64, max_samples=20, measure_speeds=True, seed=3)
    info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    _, smp = run(copy.deepcopy(info))
    speeds = {k: v.speed for k, v in smp.model.likelihood.items()}
    assert all(np.isfinite(v) and v > 1e3 for v in speeds.values()), speeds


def flat_g9(g):
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    from tests import ext_functions

    names = [str(s) for s in g["sampled"]]
    assert names == ["a", "b", "c"]
    assert [str(c) for c in g["columns"]][5:8] == ["minuslogprior", "minuslogprior__0",
                                                   "minuslogprior__ring"]
    kind = np.array([0, 0, 1], np.int32)
    lower = np.array([-1.5, -1.5, -np.inf])
    upper = np.array([1.5, 1.5, np.inf])
    likes = [LikeSpec.gaussian_mixture([0, 1, 2], [[0.1, 0.0, 0.05]], g["cov"][None],
                                       name="gaussian_mixture")]
    fm = FlatModel(names=names, prior_kind=kind, lower=lower, upper=upper, loc=np.zeros(3),
                   pscale=np.ones(3), periodic=np.zeros(3, np.int32), likes=likes,
                   ext_priors=[LikeSpec.external([0, 1], ext_functions.RING_CUDA, "ring",
                                                 name="ring")],
                   proposal_cov=np.asarray(g["proposal_cov"]), max_tries=int(g["max_tries"]))
    assert fm.columns() == [str(c) for c in g["columns"]]
    return fm


def test_external_prior_matches_reference(cuda_lib):
    """An external prior (prior.py:537-577,765-772) as a device functor: known answers and the
    chain of the unmodified reference, including its own minuslogprior__ring column."""
    g = load_golden("g9_external_prior")
    fm = flat_g9(g)
    eng = _engine(fm, 1, seed=int(g["seed"]), id0=3)
    lp, pr, ll, _ = eng.logpost(g["kat_x"])
    kat = g["kat"]            # logpost, logprior__0, logprior__ring, loglike
    finite = np.isfinite(kat[:, 0])
    assert finite.sum() >= 25 and (~finite).sum() >= 1
    np.testing.assert_array_equal(np.isfinite(lp), finite)
    np.testing.assert_allclose(lp[finite], kat[finite, 0], rtol=1e-12)
    np.testing.assert_allclose(pr[finite], kat[finite, 1] + kat[finite, 2], rtol=1e-12)
    np.testing.assert_allclose(ll[finite, 0], kat[finite, 3], rtol=1e-12, atol=1e-13)
    n = int(g["n_proposals"])
    eng.set_state(g["x0_3"][None, :])
    eng.advance(9)
    eng.advance(n - 9)
    st = eng.get_state()
    assert st["flags"][0] == 0
    ref, rows = g["rows_3"], eng.rows(0)
    assert rows.shape == ref.shape
    np.testing.assert_array_equal(rows[:, 0], ref[:, 0])
    np.testing.assert_allclose(rows, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st["x"][0], g["final_x_3"], rtol=RTOL, atol=ATOL)
    # snapshot / restore keeps the prior components of the current point
    blob = eng.export_state()
    twin = _engine(fm, 1, seed=int(g["seed"]), id0=3)
    twin.import_state(blob, eng.rows(0), np.array([len(rows)], np.int64))
    for e in (eng, twin):
        e.advance(200)
    np.testing.assert_array_equal(eng.rows(0), twin.rows(0))


def test_external_prior_through_cobaya_run(cuda_lib):
    import copy

    from tests.refenv import enable_reference

    enable_reference()
    from cobaya.run import run

    from tests import ext_functions

    info, _ = ext_functions.info_g9()
    opts = dict(info["sampler"]["mcmc"])
    opts.update(chains_per_gpu=128, max_samples=200, seed=5, Rminus1_stop=0.0)
    info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    _, smp = run(copy.deepcopy(info))
    rows = smp.products()["sample"]
    a, b = rows["a"].to_numpy(), rows["b"].to_numpy()
    np.testing.assert_allclose(rows["minuslogprior__ring"].to_numpy(),
                               -ext_functions.ring(a, b), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(rows["minuslogprior"].to_numpy(),
                               rows["minuslogprior__0"].to_numpy()
                               + rows["minuslogprior__ring"].to_numpy(), rtol=1e-12)
    w = rows["weight"].to_numpy()
    rr = np.sqrt(a * a + b * b)
    assert abs((w * rr).sum() / w.sum() - 0.5) < 0.1   # the ring holds the radius near 0.5


def test_yaml_style_lambda_strings_run_on_the_device(cuda_lib):
    """External likelihood and prior given as lambda STRINGS (the YAML form): translated to
    CUDA, checked against the Python callables by the plugin, sampled."""
    from tests.refenv import enable_reference

    enable_reference()
    from cobaya.run import run

    info = {"params": {"a": {"prior": {"min": -2, "max": 2}, "ref": 0.4, "proposal": 0.1},
                       "b": {"prior": {"min": -2, "max": 2}, "ref": 0.1, "proposal": 0.1}},
            "prior": {"ring": "lambda a,b: stats.norm.logpdf(np.sqrt(a**2+b**2), loc=0.5, scale=0.1)"},
            "likelihood": {"like1": "lambda b,a: -0.5*((a-0.1)**2+b**2)/0.09 - np.log1p(np.exp(-a))",
                           "like2": {"external": "lambda b: -0.5 * b**2 if b > -1.5 else -np.inf"}},
            "sampler": {"cobaya_b200.plugin.MCMC": {
                "chains_per_gpu": 128, "max_samples": 150, "seed": 2, "Rminus1_stop": 0.0,
                "measure_speeds": False, "covmat": np.eye(2) * 0.01, "covmat_params": ["a", "b"]}}}
    _, smp = run(info)
    rows = smp.products()["sample"]
    a, b = rows["a"].to_numpy(), rows["b"].to_numpy()
    rr = np.sqrt(a * a + b * b)
    np.testing.assert_allclose(rows["minuslogprior__ring"].to_numpy(),
                               0.5 * ((rr - 0.5) / 0.1) ** 2 + np.log(0.1)
                               + 0.5 * np.log(2 * np.pi), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(rows["chi2__like1"].to_numpy(),
                               ((a - 0.1) ** 2 + b * b) / 0.09 + 2 * np.log1p(np.exp(-a)),
                               rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(rows["chi2__like2"].to_numpy(), b * b, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rows["chi2"].to_numpy(), rows["chi2__like1"].to_numpy()
                               + rows["chi2__like2"].to_numpy(), rtol=1e-12)


def test_external_function_under_dragging_equals_the_builtin_kernel(cuda_lib, route):
    """BASELINE configs[3] as stated -- Rosenbrock as an EXTERNAL likelihood, dragging: the
    split-launch route (k_extd_*) walks like k_step_general on the built-in Rosenbrock (same
    draws; the two sums differ in the order of the terms only)."""
    from tests import ext_functions

    builtin, ext, start = ext_functions.rosenbrock_pair()
    C, n = 48, 70
    x0 = start(C, 0)
    a, b = _engine(builtin, C, 31), _engine(ext, C, 31)
    a.set_kernel_policy(1)   # general kernel
    for e in (a, b):
        e.set_state(x0)
        e.advance(7)
        e.advance(n - 7)
    sa, sb = a.get_state(), b.get_state()
    assert not sb["flags"].any() and not sa["flags"].any()
    np.testing.assert_array_equal(sa["n_rows"], sb["n_rows"])
    np.testing.assert_array_equal(sa["weight"], sb["weight"])
    assert sa["n_rows"].sum() > C * 5
    np.testing.assert_allclose(sa["x"], sb["x"], rtol=1e-9, atol=1e-12)
    ra, ca = a.rows_bulk()
    rb, cb = b.rows_bulk()
    np.testing.assert_array_equal(ra[:, 0], rb[:, 0])
    np.testing.assert_allclose(ra, rb, rtol=1e-9, atol=1e-11)
    _check_route(b, route)
    # and like the fused dragging kernel (k_step_drag) on the built-in function
    c = _engine(builtin, C, 31)
    c.set_state(x0)
    c.advance(n)
    assert c.last_step_kernel() == 1
    rc_, cc = c.rows_bulk()
    np.testing.assert_array_equal(cc, cb)
    np.testing.assert_allclose(rc_, rb, rtol=1e-9, atol=1e-11)


def test_external_dragging_through_cobaya_run(cuda_lib):
    from tests.refenv import enable_reference

    enable_reference()
    from cobaya.run import run

    names = ["a", "b", "c", "d"]
    info = {"params": {p: {"prior": {"min": -3, "max": 3}, "ref": 0.9, "proposal": 0.1}
                       for p in names},
            "likelihood": {"rosen": "lambda a, b, c, d: -(100*(b-a**2)**2 + (1-a)**2 + "
                                    "100*(c-b**2)**2 + (1-b)**2 + 100*(d-c**2)**2 + (1-c)**2)/20"},
            "sampler": {"cobaya_b200.plugin.MCMC": {
                "chains_per_gpu": 64, "max_samples": 60, "seed": 4, "Rminus1_stop": 0.0,
                "drag": True, "blocking": [[1, ["a", "b"]], [4, ["c", "d"]]],
                "measure_speeds": False, "covmat": np.eye(4) * 0.01, "covmat_params": names}}}
    _, smp = run(info)
    assert smp._fm.drag and smp._fm.drag_interp_steps >= 1
    rows = smp.products()["sample"]
    a, b, c, d = (rows[p].to_numpy() for p in names)
    want = (100 * (b - a**2) ** 2 + (1 - a) ** 2 + 100 * (c - b**2) ** 2 + (1 - b) ** 2
            + 100 * (d - c**2) ** 2 + (1 - c) ** 2) / 10
    np.testing.assert_allclose(rows["chi2__rosen"].to_numpy(), want, rtol=1e-9, atol=1e-9)
    assert len(rows) >= 64 * 60
