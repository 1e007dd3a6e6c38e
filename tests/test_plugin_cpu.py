"""
CPU tests of the drop-in boundary: the plugin class is accepted by
the reference's own component machinery, its host-side setup (blocking, thinning, covmat
-> transforms, start points) equals the reference's MCMC.initialize on the same input, the
lowered model evaluates like Model.logposterior, and -- without a GPU -- run() fails loudly
instead of falling back to a CPU path.
"""

import copy

import numpy as np
import pytest

from tests.refenv import enable_reference


def _infos():
    enable_reference()
    from oracle import make_golden as mg

    out = {}
    for name, fn in (("g2", mg.info_g2), ("g3", mg.info_g3), ("g4", mg.info_g4)):
        try:
            out[name] = fn()[0]
        except FileNotFoundError:  # g4 reads /root/reference/tests: only in the build box
            continue
    return out


def _pair(info, **extra):
    from cobaya.model import get_model
    from cobaya.sampler import get_sampler

    ref_info = copy.deepcopy(info)
    ref_model = get_model(ref_info)
    ref = get_sampler(ref_info["sampler"], ref_model)
    b2_info = copy.deepcopy(info)
    opts = dict(b2_info["sampler"]["mcmc"], chains_per_gpu=3, **extra)
    b2_info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    model = get_model(b2_info)
    mine = get_sampler(b2_info["sampler"], model)
    return ref_model, ref, model, mine


@pytest.mark.parametrize("case", ["g2", "g3", "g4"])
def test_initialize_matches_reference(case):
    infos = _infos()
    if case not in infos:
        pytest.skip("needs /root/reference/tests")
    ref_model, ref, model, mine = _pair(infos[case])
    assert [list(b) for b in mine.blocks] == [list(b) for b in ref.blocks]
    assert list(mine.oversampling_factors) == list(ref.oversampling_factors)
    assert mine.cycle_length == ref.cycle_length
    assert mine.current_point.output_thin == ref.current_point.output_thin
    assert bool(mine.drag) == bool(ref.drag)
    if ref.drag:
        assert mine.drag_interp_steps == ref.drag_interp_steps
        assert mine._fm.last_slow == ref.proposer.i_last_slow_block
    np.testing.assert_array_equal(mine.proposer.i_of_j, ref.proposer.i_of_j)
    np.testing.assert_allclose(mine.proposer.get_covariance(), ref.proposer.get_covariance(),
                               rtol=0, atol=0)
    for a, b in zip(mine.proposer.transform, ref.proposer.transform):
        np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-18)
    assert mine._fm.columns() == list(ref.collection.columns)
    assert mine._x0.shape == (3, model.prior.d())
    for x in mine._x0:
        assert np.isfinite(model.logposterior(x).logpost)


def test_lowered_model_evaluates_like_reference_logposterior():
    """FlatModel (lowered from the live reference objects) through the C oracle vs
    Model.logposterior, including the -inf branch (model.py:650)."""
    from oracle import oracle as orc

    infos = _infos()
    ref_model, ref, model, mine = _pair(infos["g2"])
    om = orc.OracleModel(mine._fm)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.45, 0.45, (20, 5))
    pts[3, 0] = 2.0  # outside the prior
    for p in pts:
        r = model.logposterior(p)
        v, lp, ll, der = om.logpost(p)
        if np.isfinite(r.logpost):
            np.testing.assert_allclose(v, r.logpost, rtol=1e-11)
            np.testing.assert_allclose(lp, r.logprior, rtol=1e-12)
            np.testing.assert_allclose(der, r.derived, rtol=1e-9, atol=1e-12)
        else:
            assert v == -np.inf and len(r.loglikes) == 0


@pytest.mark.parametrize("case", ["g5", "g6"])
def test_lowering_of_scipy_priors_and_more_likelihoods(case):
    """the scipy.stats 1-D priors and the ``gaussian`` / ``one``
    likelihoods are lowered from the live reference objects; the lowered model (through
    the C oracle) evaluates like Model.logposterior."""
    from oracle import oracle as orc

    enable_reference()
    from oracle import make_golden as mg

    info = (mg.info_g5 if case == "g5" else mg.info_g6)()[0]
    ref_model, ref, model, mine = _pair(info)
    fm = mine._fm
    if case == "g5":
        assert sorted(set(int(k) for k in fm.prior_kind)) == list(range(11))
    else:
        assert [lk.name for lk in fm.likes] == ["gaussian", "one"]
        assert fm.columns()[-2:] == ["chi2__gaussian", "chi2__one"]
    om = orc.OracleModel(fm)
    rng = np.random.default_rng(1)
    center = np.array([model.prior.reference(random_state=rng) for _ in range(1)])[0]
    n_in = 0
    for _ in range(60):
        p = center + 0.3 * rng.standard_normal(len(center))
        r = model.logposterior(p)
        v, lp, ll, der = om.logpost(p)
        if np.isfinite(r.logpost):
            n_in += 1
            np.testing.assert_allclose(lp, r.logprior, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(ll, r.loglikes, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(v, r.logpost, rtol=1e-10)
        else:
            assert v == -np.inf
    assert n_in > 5


def test_unsupported_prior_is_refused():
    enable_reference()
    from cobaya.log import LoggedError
    from cobaya.model import get_model
    from cobaya.sampler import get_sampler

    from oracle import make_golden as mg

    info = mg.info_g6()[0]
    info["params"]["q1"]["prior"] = {"dist": "weibull_min", "c": 1.5, "loc": -1, "scale": 1}
    info["sampler"] = {"cobaya_b200.plugin.MCMC": dict(info["sampler"]["mcmc"],
                                                       chains_per_gpu=2)}
    with pytest.raises(LoggedError, match="recognised set"):
        get_sampler(info["sampler"], get_model(info))


def test_unknown_option_is_rejected_and_engine_keys_are_accepted():
    enable_reference()
    from cobaya.input import update_info
    from cobaya.log import LoggedError

    from oracle import make_golden as mg

    info = mg.info_g2()[0]
    info["sampler"] = {"cobaya_b200.plugin.MCMC": dict(info["sampler"]["mcmc"],
                                                       chains_per_gpu=16, rows_per_chain=100)}
    upd = update_info(copy.deepcopy(info))
    opts = upd["sampler"]["cobaya_b200.plugin.MCMC"]
    assert opts["chains_per_gpu"] == 16 and opts["learn_every"] == "40d"
    assert opts["Rminus1_stop"] == 0.01 and opts["proposal_scale"] == 2.4
    bad = copy.deepcopy(info)
    bad["sampler"]["cobaya_b200.plugin.MCMC"]["no_such_option"] = 1
    # the rejection path formats fuzzy suggestions with `rapidfuzz` (tools.py:859), which is
    # not installed in this image: either way the unknown key is refused
    with pytest.raises((LoggedError, ModuleNotFoundError)):
        update_info(bad)


def test_every_reference_mcmc_option_is_mirrored():
    enable_reference()
    from cobaya.samplers.mcmc import MCMC as RefMCMC

    from cobaya_b200.plugin import MCMC

    ref = RefMCMC.get_defaults()
    mine = MCMC.get_defaults()
    for k, v in ref.items():
        assert k in mine, k
        if isinstance(v, float) and np.isinf(v):
            assert np.isinf(mine[k])
        else:
            assert mine[k] == v, (k, mine[k], v)


def test_install_as_mcmc_makes_sampler_mcmc_resolve_to_the_engine():
    enable_reference()
    import sys

    from cobaya.sampler import get_sampler_name_and_class

    import cobaya_b200.plugin as plugin

    saved = sys.modules.get("cobaya.samplers.mcmc")
    try:
        plugin.install_as_mcmc()
        name, cls = get_sampler_name_and_class({"mcmc": None})
        assert name == "mcmc" and cls is plugin.MCMC
    finally:
        if saved is not None:
            sys.modules["cobaya.samplers.mcmc"] = saved


def test_unsupported_model_raises_instead_of_falling_back():
    enable_reference()
    from cobaya.log import LoggedError
    from cobaya.model import get_model
    from cobaya.sampler import get_sampler

    info = {"likelihood": {"ext": lambda a: -0.5 * a**2},
            "params": {"a": {"prior": {"min": -3, "max": 3}, "proposal": 0.5}},
            "sampler": {"cobaya_b200.plugin.MCMC": {"chains_per_gpu": 2,
                                                    "measure_speeds": False}}}
    model = get_model(info)
    with pytest.raises(LoggedError, match="cannot be evaluated on the device"):
        get_sampler(info["sampler"], model)


def test_run_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    enable_reference()
    from cobaya.log import LoggedError

    infos = _infos()
    _, _, _, mine = _pair(infos["g2"])
    with pytest.raises(LoggedError, match="no CPU fallback"):
        mine.run()


def test_products_and_output_files_roundtrip(tmp_path):
    """SURVEY 8f row 1 (output side): the plugin materialises real SampleCollection products
    and, with an `output` prefix, writes chain / covmat / checkpoint files that the
    reference's own loader reads back (rows here come from the C oracle standing in for the
    device, the engine itself needs a GPU)."""
    enable_reference()
    from cobaya.model import get_model
    from cobaya.output import get_output, load_samples
    from cobaya.sampler import get_sampler

    from oracle import oracle as orc

    infos = _infos()
    info = copy.deepcopy(infos["g2"])
    opts = dict(info["sampler"]["mcmc"], chains_per_gpu=3)
    info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    prefix = str(tmp_path / "run")
    out = get_output(prefix=prefix, resume=False, force=True)
    model = get_model(info)
    s = get_sampler(info["sampler"], model, output=out)

    class FakeEnsemble:  # rows with the engine's layout, produced by the oracle
        def __init__(self, fm, x0):
            om = orc.OracleModel(fm)
            self.rows = [orc.OracleChain(om, 7, c, x0[c], burn_in=2).advance(600)[1]
                         for c in range(len(x0))]
            self.n_chains_local = len(x0)
            self.engine = self
            self.visible = 0.4  # fraction of every chain "sampled so far"

        def rows_bulk(self, first=None, chains=None):  # Engine.rows_bulk
            first = np.zeros(len(self.rows), np.int64) if first is None else first
            part = [r[f: int(self.visible * len(r))] for r, f in zip(self.rows, first)]
            return np.concatenate(part), np.array([len(q) for q in part], np.int64)

        def samples(self, chains=None, skip_samples=0.0):
            k = lambda r: int(skip_samples * len(r)) if 0 < skip_samples < 1 else int(skip_samples)
            return np.concatenate([r[k(r):] for r in self.rows])

    s._ens = FakeEnsemble(s._fm, s._x0)
    s._row_cursor = np.zeros(3, np.int64)
    s._drain_rows()                    # a timed output in the middle of the run ...
    n_first = len(s.collection)
    assert 0 < n_first == sum(int(0.4 * len(r)) for r in s._ens.rows)
    assert len(load_samples(prefix, skip=0, combined=True)) == n_first  # ... is on disk
    s._ens.visible = 1.0
    s._drain_rows()                    # ... and the rest at the end: appended, not rewritten
    s.write_checkpoint()
    col = s.products()["sample"]
    n = sum(len(r) for r in s._ens.rows)
    assert len(col) == n and list(col.columns) == s._fm.columns()
    # layout: one segment per output, chains in order inside a segment
    assert [list(c) for c in s._segments] == [
        [int(0.4 * len(r)) for r in s._ens.rows],
        [len(r) - int(0.4 * len(r)) for r in s._ens.rows]]
    np.testing.assert_array_equal(col.data.to_numpy()[: int(0.4 * len(s._ens.rows[0]))],
                                  s._ens.rows[0][: int(0.4 * len(s._ens.rows[0]))])
    raw = np.fromfile(s.rows_filename()).reshape(n, -1)   # binary twin of the chain file
    np.testing.assert_array_equal(raw, col.data.to_numpy())
    # the g2 case samples at temperature 2: compare the statistics of the tempered sample
    np.testing.assert_allclose(col.mean(tempered=True), np.average(
        np.concatenate(s._ens.rows)[:, 2:7], axis=0, weights=np.concatenate(s._ens.rows)[:, 0]))
    half = s.products(skip_samples=0.5)["sample"]
    assert 0 < len(half) < n
    import os

    files = sorted(os.listdir(tmp_path))
    assert "run.1.txt" in files and "run.covmat" in files and "run.checkpoint" in files
    back = load_samples(prefix, skip=0, combined=True)
    # the chain file keeps every row; skip_samples only affects the returned copy
    assert len(back) == n and len(s.collection) == n
    np.testing.assert_allclose(back[["p0", "p1"]].to_numpy(), col[["p0", "p1"]].to_numpy(),
                               rtol=1e-7)
    np.testing.assert_allclose(half[["p0", "p1"]].to_numpy(),
                               s._ens.samples(skip_samples=0.5)[:, 2:4], rtol=0)
    # .progress: the reference's format (mcmc.py:163-181,1067-1077), read by its own loader
    import datetime

    from cobaya.tools import load_DataFrame

    for i, (N, acc, r1, rcl) in enumerate([(1200, 0.31, 0.8, None), (2400, 0.29, 0.05, 0.4)], 1):
        s.progress.loc[i] = [N, datetime.datetime.now().isoformat(), acc, r1, rcl]
    s.write_checkpoint()
    assert "run.progress" in os.listdir(tmp_path)
    with open(s.progress_filename()) as f:
        lines = f.readlines()
    assert lines[0].startswith("#") and len(lines) == 3
    prog = load_DataFrame(s.progress_filename())
    assert list(prog.columns) == ["N", "timestamp", "acceptance_rate", "Rminus1", "Rminus1_cl"]
    np.testing.assert_allclose(prog["N"], [1200, 2400])
    np.testing.assert_allclose(prog["Rminus1"], [0.8, 0.05])
    assert np.isnan(prog["Rminus1_cl"][0]) and prog["Rminus1_cl"][1] == 0.4


def test_reference_minimize_and_post_consume_the_engine_output(tmp_path):
    """SURVEY 8f row 4 (consumers): the files the plugin writes -- per-process chain file,
    ``.covmat``, ``updated.yaml`` -- are what the reference's own ``minimize`` sampler
    (start from the best sample, step sizes from the learned covmat:
    samplers/minimize/minimize.py:170-250,435) and ``post`` (importance re-weighting over
    the stored rows: post.py:279-600) read.  Chains come from the oracle-backed test
    engine (no GPU here); everything downstream is the unmodified reference."""
    enable_reference()
    from cobaya.post import post
    from cobaya.run import run

    import cobaya_b200.plugin as plugin
    from tests.oracle_engine import OracleEngine

    g = __import__("tests.util", fromlist=["load_golden"]).load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)
    prefix = str(tmp_path / "run")
    info = {
        "likelihood": {"gaussian_mixture": {"means": [mean], "covs": [cov],
                                            "input_params_prefix": "a_",
                                            "output_params_prefix": "", "derived": False}},
        "params": {f"a__{i}": {"prior": {"min": -1, "max": 1}} for i in range(3)},
        "sampler": {"cobaya_b200.plugin.MCMC": {
            "covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
            "burn_in": 5, "max_tries": 3000, "learn_proposal_Rminus1_max": 30,
            "Rminus1_stop": 1e-9, "measure_speeds": False, "seed": 3, "chains_per_gpu": 6,
            "max_samples": 700, "learn_every": "20d"}},
        "output": prefix,
    }
    plugin.MCMC._engine_factory = OracleEngine
    try:
        _, smp = run(copy.deepcopy(info), force=True)
    finally:
        plugin.MCMC._engine_factory = None
    learned = smp.proposer.get_covariance()
    # ---- minimize: step sizes from the covmat file the plugin wrote.  (Sharing the output
    # prefix, so that it also starts from the chain's best point, runs into a bug of the
    # reference itself: minimize.py:197 passes `concatenate=` to Output.load_collections,
    # which output.py:374 does not accept.)
    info_min = copy.deepcopy(info)
    del info_min["output"]
    info_min["sampler"] = {"minimize": {"method": "scipy", "best_of": 1, "ignore_prior": False,
                                        "covmat": prefix + ".covmat", "seed": 1}}
    _, mini = run(info_min)
    x_min = np.array([mini.products()["minimum"][f"a__{i}"] for i in range(3)])
    np.testing.assert_allclose(x_min, mean, atol=2e-4)
    # the step scales of the minimizer come from the covmat the engine wrote
    np.testing.assert_allclose(mini._scales, 1 / np.sqrt(np.diag(np.linalg.inv(learned))),
                               rtol=1e-5)
    # ---- post: re-weight the stored sample to a shifted target
    shift = np.array([0.02, -0.01, 0.0])
    info_post = {"output": prefix, "post": {
        "suffix": "shifted", "skip": 0.3,
        "remove": {"likelihood": {"gaussian_mixture": None}},
        "add": {"likelihood": {"gaussian_mixture": {
            "means": [mean + shift], "covs": [cov], "input_params_prefix": "a_",
            "output_params_prefix": "", "derived": False}}}}}
    _, res = post(info_post)
    col = res["sample"]
    col = col[0] if isinstance(col, list) else col
    before = smp.products(skip_samples=0.3)["sample"]
    assert len(col) > 100
    moved = np.array(col.mean()[:3]) - np.array(before.mean()[:3])
    # importance re-weighting moves the mean towards the new target
    assert np.dot(moved, shift) > 0.5 * np.dot(shift, shift)


def test_vectorised_start_points_follow_ref_and_prior():
    """More than 64 chains per process: the start points come from the vectorised
    restatement of Prior.reference / Model.get_valid_point (prior.py:866-961,
    model.py:707-754): ``ref`` pdf where given, fixed ``ref`` values kept, the prior where no
    ``ref`` is given, every point inside the prior's support."""
    enable_reference()
    from cobaya.model import get_model
    from cobaya.sampler import get_sampler

    cov = np.diag([0.01, 0.04, 0.02, 0.01]) ** 2
    info = {
        "likelihood": {"gaussian_mixture": {"means": [np.zeros(4)], "covs": [cov],
                                            "input_params": ["a", "b", "c", "d"],
                                            "derived": False}},
        "params": {
            "a": {"prior": {"min": -1, "max": 1}, "ref": {"dist": "norm", "loc": 0.2, "scale": 0.01}},
            "b": {"prior": {"min": 0, "max": 0.5}, "ref": 0.25},
            "c": {"prior": {"min": -0.1, "max": 0.1}},                     # no ref: from the prior
            "d": {"prior": {"dist": "norm", "loc": 0, "scale": 0.1},
                  "ref": {"dist": "norm", "loc": 0, "scale": 3.0}}},
        "sampler": {"cobaya_b200.plugin.MCMC": {"measure_speeds": False, "chains_per_gpu": 4000,
                                                "seed": 2}},
    }
    model = get_model(info)
    s = get_sampler(info["sampler"], model)
    x = s._x0
    assert x.shape == (4000, 4) and np.all(np.isfinite(x))
    assert abs(x[:, 0].mean() - 0.2) < 1e-3 and abs(x[:, 0].std() - 0.01) < 1e-3
    assert np.all(x[:, 1] == 0.25)
    assert x[:, 2].min() >= -0.1 and x[:, 2].max() <= 0.1 and abs(x[:, 2].std() - 0.2 / 12 ** 0.5) < 4e-3
    assert abs(x[:, 3].std() - 3.0) < 0.2
    for row in x[:20]:
        assert np.isfinite(model.logposterior(row).logpost)


def test_external_function_sources_compile_and_lower_without_a_gpu():
    """NVRTC needs no device: the CUDA twin of an external likelihood is compiled (and its
    errors reported) on the CPU; the lowering turns the reference's
    LikelihoodExternalFunction into an external LikeSpec."""
    enable_reference()
    from cobaya.model import get_model

    from tests import ext_functions
    from cobaya_b200.flatmodel import LIKE_EXTERNAL
    from cobaya_b200.functor import DeviceFunctionError, check_source, device_function
    from cobaya_b200.lowering import UnsupportedModelError, lower_likelihoods

    assert check_source(ext_functions.BANANA_CUDA, dim=2) == ""
    with pytest.raises(DeviceFunctionError, match="undefined"):
        check_source('extern "C" __device__ double f(const double *p, int n) { return q; }')
    with pytest.raises(DeviceFunctionError, match="exactly one"):
        device_function("double f(double x) { return x; }")
    info, _ = ext_functions.info_g8()
    info.pop("sampler")
    model = get_model(info)
    sampled = list(model.parameterization.sampled_params())
    likes = lower_likelihoods(model, sampled)
    assert [lk.kind for lk in likes][0] == LIKE_EXTERNAL and likes[0].fn_name == "banana"
    assert list(likes[0].idx) == [sampled.index("a"), sampled.index("b")]
    # a plain Python callable has no device twin: refused, never a CPU fallback
    info["likelihood"]["banana"] = {"external": lambda a, b: -a * a - b * b}
    with pytest.raises(UnsupportedModelError, match="device_function"):
        lower_likelihoods(get_model(info), sampled)
    # external priors (prior.py:537-577) the same way; the row gets their own column
    from cobaya_b200.lowering import lower_model

    info9, _ = ext_functions.info_g9()
    info9.pop("sampler")
    fm = lower_model(get_model(info9), proposal_cov=np.eye(3))
    assert [ep.name for ep in fm.ext_priors] == ["ring"] and list(fm.ext_priors[0].idx) == [0, 1]
    assert fm.columns()[5:8] == ["minuslogprior", "minuslogprior__0", "minuslogprior__ring"]
    assert fm.row_width == len(fm.columns()) == 10
    info9["prior"] = {"ring": lambda a, b: -(a * a + b * b)}
    with pytest.raises(UnsupportedModelError, match="External prior 'ring'"):
        lower_model(get_model(info9), proposal_cov=np.eye(3))


def test_lambda_strings_are_translated_to_cuda_on_the_cpu():
    """YAML-style external functions (lambda strings, tools.py:344-384) get their CUDA twin
    automatically; what the small grammar does not cover is refused with a reason."""
    enable_reference()
    from cobaya.model import get_model

    from cobaya_b200.functor import DeviceFunctionError, check_source, cuda_from_lambda
    from cobaya_b200.lowering import UnsupportedModelError, lower_model

    src, entry, args = cuda_from_lambda(
        "lambda a, b: stats.norm.logpdf(a - b**2, loc=0, scale=0.1) - np.log1p(a*a) "
        "+ (0 if a > -1 and b < 2 else -np.inf) + np.minimum(a, 0.5) * np.pi", "my like")
    assert args == ["a", "b"] and entry == "cb2_fn_my_like"
    assert check_source(src, entry, dim=2) == ""
    for bad in ("lambda a: [a]", "lambda a: a.real", "lambda a, _self: a", "lambda *a: 1.0",
                "lambda a, b: a and b", "lambda a: a // 2", "lambda a: a % 2",
                "lambda a: scipy.special.gamma(a)", "lambda a: q + a", "import os"):
        with pytest.raises(DeviceFunctionError):
            cuda_from_lambda(bad, "f")
    info = {"params": {"a": {"prior": {"min": -1, "max": 1}},
                       "b": {"prior": {"min": -1, "max": 1}}},
            "prior": {"ring": "lambda a,b: stats.norm.logpdf(np.sqrt(a**2+b**2), loc=0.5, scale=0.1)"},
            "likelihood": {"like1": "lambda b,a: -0.5*(a**2+b**2)/0.04",
                           "like2": {"external": "lambda b: -b**2"}}}
    fm = lower_model(get_model(info), proposal_cov=np.eye(2))
    assert [list(lk.idx) for lk in fm.likes] == [[1, 0], [1]]   # p[] follows the signature
    assert fm.columns()[4:] == ["minuslogprior", "minuslogprior__0", "minuslogprior__ring",
                                "chi2", "chi2__like1", "chi2__like2"]
    for lk in fm.likes + fm.ext_priors:
        assert check_source(lk.source, lk.fn_name, dim=lk.dim) == ""
    info["likelihood"]["like1"] = "lambda a, b: float(np.linalg.norm([a, b]))"
    with pytest.raises(UnsupportedModelError, match="cannot be evaluated on the device"):
        lower_model(get_model(info), proposal_cov=np.eye(2))


def test_the_fused_external_kernel_compiles_without_a_gpu():
    """The general step kernel is recompiled at run time with the user's likelihood functions
    inlined (csrc/ext_functor.inl, from the copy of the kernel sources embedded in the library):
    the NVRTC compile needs no device."""
    import ctypes as C

    from cobaya_b200 import _cabi
    from cobaya_b200.functor import cuda_from_lambda
    from cobaya_b200.problems import ROSENBROCK_CUDA

    lib = _cabi.load()
    log = C.create_string_buffer(1 << 14)
    assert lib.cb2_check_external_fused(ROSENBROCK_CUDA.encode(), b"rosenbrock_ext", 30, log,
                                        len(log)) == 0, log.value.decode()
    src, entry, _ = cuda_from_lambda("lambda a, b: -a**2 - np.log1p(b * b)", "f")
    assert lib.cb2_check_external_fused(src.encode(), entry.encode(), 2, log, len(log)) == 0
    bad = 'extern "C" __device__ double f(const double *p, int n) { return nope; }'
    assert lib.cb2_check_external_fused(bad.encode(), b"f", 2, log, len(log)) == -7
    assert "nope" in log.value.decode()


LAMBDAS = [
    "lambda a, b: stats.norm.logpdf(a - b**2, loc=0.1, scale=0.3) - np.log1p(a*a) + 0.5*np.pi",
    "lambda a, b: -0.5*((a-0.1)**2+b**2)/0.09 - np.log1p(np.exp(-a)) if b > -1.5 else -np.inf",
    "lambda a, b: np.minimum(a, 0.3) * np.maximum(b, -0.2) - abs(a - b) + np.arctan2(a, 1 + b*b)",
    "lambda a, b: stats.uniform.logpdf(a, loc=-2, scale=4) + np.sqrt(a*a + b*b + 1e-3)**3 / 7",
    "lambda b, a: -(100*(b-a**2)**2 + (1-a)**2)/20 + np.tanh(a) * np.cos(b) - np.power(1.5, a)",
    "lambda a, b: (a if (a > 0 and b > 0) or a < -1 else -a) + math.erf(b) - np.log10(2 + a*a)",
    "lambda a, b: -50 * (a > 1.2) + (a - b)**2 * (b <= 0.4) - (1 if a else 2)",
]


@pytest.mark.parametrize("text", LAMBDAS)
def test_translated_lambdas_compute_what_python_computes(text, tmp_path):
    """functor.cuda_from_lambda is a small compiler: its output is plain C arithmetic, so it
    is compiled here with g++ as host code (the CUDA qualifiers defined away) and compared with
    the Python callable the reference would evaluate (tools.py:344-384: ``np`` and ``stats``
    in scope) on random points."""
    import ctypes as C
    import math
    import subprocess

    from scipy import stats

    from cobaya_b200.functor import cuda_from_lambda

    src, entry, args = cuda_from_lambda(text, "f")
    prelude = ("#include <cmath>\n#include <cstring>\n#define __device__\n"
               "#define __forceinline__ inline\n"
               "static inline double __longlong_as_double(long long x) "
               "{ double d; std::memcpy(&d, &x, 8); return d; }\n")
    cfile, so = tmp_path / "f.cpp", tmp_path / "f.so"
    cfile.write_text(prelude + src)
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", str(so), str(cfile)], check=True)
    fn = getattr(C.CDLL(str(so)), entry)
    fn.restype = C.c_double
    fn.argtypes = [C.POINTER(C.c_double), C.c_int]
    py = eval(text, {"np": np, "stats": stats, "math": math})
    rng = np.random.default_rng(7)
    pts = rng.uniform(-1.9, 1.9, (300, len(args)))
    for p in pts:
        want = float(py(*p))
        got = fn((C.c_double * len(p))(*p), len(p))
        if np.isfinite(want):
            assert got == pytest.approx(want, rel=1e-12, abs=1e-13), (text, p)
        else:
            assert got == want


def test_external_functions_through_the_plugin_on_the_cpu_test_engine():
    """The plugin's handling of external functions (lowering of the callable's CUDA twin or
    lambda string, the start-point check of the twin against the Python callable, rows with
    the function's own chi2 column) with the oracle-backed test engine, which evaluates the
    SAME CUDA source compiled as host code: no GPU needed for the host logic."""
    enable_reference()
    from cobaya.log import LoggedError
    from cobaya.run import run

    import cobaya_b200.plugin as plugin
    from cobaya_b200.functor import device_function
    from tests import ext_functions
    from tests.oracle_engine import OracleEngine

    info, _ = ext_functions.info_g8()
    opts = dict(info["sampler"]["mcmc"])
    opts.update(chains_per_gpu=6, max_samples=60, learn_proposal=False, seed=3,
                Rminus1_stop=1e-9)
    info["sampler"] = {"cobaya_b200.plugin.MCMC": opts}
    # a second external component given as a lambda string (the YAML form)
    info["likelihood"]["soft"] = "lambda c: -0.5 * c**2 / 4.0 - np.log1p(c*c)"
    plugin.MCMC._engine_factory = OracleEngine
    try:
        _, smp = run(copy.deepcopy(info))
        rows = smp.products()["sample"]
        a, b, c = (rows[p].to_numpy() for p in "abc")
        np.testing.assert_allclose(rows["chi2__banana"].to_numpy(),
                                   -2 * ext_functions.banana(a, b), rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(rows["chi2__soft"].to_numpy(),
                                   c**2 / 4.0 + 2 * np.log1p(c * c), rtol=1e-10, atol=1e-11)
        assert len(rows) >= 6 * 60
        # a twin that computes something else is refused before sampling
        bad = device_function(ext_functions.BANANA_CUDA.replace("0.09", "0.08"))(
            lambda a, b: ext_functions.banana(a, b))
        info2 = copy.deepcopy(info)
        info2["likelihood"]["banana"] = {"external": bad}
        with pytest.raises(LoggedError, match="disagrees"):
            run(info2)
        # an external prior (own minuslogprior__ring column) through the same machinery
        info9, _ = ext_functions.info_g9()
        opts9 = dict(info9["sampler"]["mcmc"])
        opts9.update(chains_per_gpu=5, max_samples=50, seed=4, Rminus1_stop=1e-9)
        info9["sampler"] = {"cobaya_b200.plugin.MCMC": opts9}
        _, smp9 = run(info9)
        rows9 = smp9.products()["sample"]
        a9, b9 = rows9["a"].to_numpy(), rows9["b"].to_numpy()
        np.testing.assert_allclose(rows9["minuslogprior__ring"].to_numpy(),
                                   -ext_functions.ring(a9, b9), rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(rows9["minuslogprior"].to_numpy(),
                                   rows9["minuslogprior__0"].to_numpy()
                                   + rows9["minuslogprior__ring"].to_numpy(), rtol=1e-12)
    finally:
        plugin.MCMC._engine_factory = None
