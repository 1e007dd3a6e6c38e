"""
Golden-vector generator (TEST INFRASTRUCTURE; runs only in the build container).

Drives the UNMODIFIED reference (cobaya 3.6.2, imported from baseline/_ref or
/root/reference, plus the getdist import shim in oracle/shims) and records its
outputs under tests/golden/.  The reference's classes -- ``MCMC``,
``BlockedProposer``, ``RandDirectionProposer``, ``CyclicIndexRandomizer``,
numba ``_rvs``, ``Model.logposterior``, ``SampleCollection`` -- run as shipped;
the only substitution is the ``random_state`` object they draw from, which
returns the engine's counter-based Philox draws (same distributions as the
numpy calls it replaces) so that the chain the reference produces can be
compared value-for-value with the oracle and the CUDA engine.

Usage:  python oracle/make_golden.py [fixture ...]   (rewrites tests/golden/*.npz)
"""

from __future__ import annotations

import linecache
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(cand, "cobaya")):
        sys.path.insert(0, cand)
        break

from oracle import oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------------
class Ctx:
    def __init__(self, seed, chain_id):
        self.seed, self.chain_id = seed, chain_id
        self.t = 0
        self.sub = 0


class _CyclerRS:
    """random_state of one CyclicIndexRandomizer (proposal.py:54)."""

    def __init__(self, ctx, which):
        self.ctx, self.which, self.cycle = ctx, which, -1

    def permutation(self, sorted_indices):
        self.cycle += 1
        return orc.permutation(np.asarray(sorted_indices), self.ctx.seed,
                               self.ctx.chain_id, self.which, self.cycle)


class _BlockRS:
    """random_state of one RandDirectionProposer / RandProposer1D."""

    def __init__(self, ctx, block, n):
        self.ctx, self.block, self.n, self.epoch = ctx, block, n, -1

    def standard_normal(self, size):  # functions.py:36
        self.epoch += 1
        xx = orc.basis_normals(self.n, self.ctx.seed, self.ctx.chain_id, self.block,
                               self.epoch)
        assert len(xx) == size
        return xx

    def _words(self, tag):
        t = self.ctx.t
        return orc.philox4x32(
            (self.ctx.seed & 0xFFFFFFFF, self.ctx.seed >> 32),
            (t & 0xFFFFFFFF, t >> 32, self.ctx.chain_id, tag | (self.ctx.sub << 8)))

    def uniform(self):  # proposal.py:79
        return (float(self._words(0)[0]) + 0.5) * 2.0**-32

    def integers(self, n):  # proposal.py:90
        assert n == 2
        return int(self._words(0)[1] & 1)

    def standard_exponential(self):  # proposal.py:80
        r, _ = orc.radial(max(self.n, 2), self.ctx.seed, self.ctx.chain_id, self.ctx.t,
                          self.ctx.sub)
        # only called on the u<0.33 branch, where orc_radial returns -log(u_r)
        return r

    def chisquare(self, df):  # proposal.py:82
        r, _ = orc.radial(self.n, self.ctx.seed, self.ctx.chain_id, self.ctx.t,
                          self.ctx.sub)
        return r * r


class _AcceptRS:
    """sampler._rng as used by MCMC.metropolis_accept (mcmc.py:683)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def standard_exponential(self):
        caller = sys._getframe(2)
        sub = 0
        if caller.f_code.co_name == "get_new_sample_dragging":
            line = linecache.getline(caller.f_code.co_filename, caller.f_lineno)
            if "accept_drag" in line:
                sub = self.ctx.sub  # drag step i
        return orc.accept_exp(self.ctx.seed, self.ctx.chain_id, self.ctx.t, sub)


def install_philox(sampler, seed, chain_id):
    ctx = Ctx(seed, chain_id)
    pr = sampler.proposer
    pr.block_cycler.random_state = _CyclerRS(ctx, 0)
    pr.block_cycler_slow.random_state = _CyclerRS(ctx, 1)
    pr.block_cycler_fast.random_state = _CyclerRS(ctx, 2)
    for b, bp in enumerate(pr.proposer):
        bp.random_state = _BlockRS(ctx, b, bp.n)
    sampler._rng = _AcceptRS(ctx)
    slow, fast, main = pr.get_proposal_slow, pr.get_proposal_fast, pr.get_proposal

    def _slow(P):
        ctx.sub = 0
        return slow(P)

    def _fast(P):
        ctx.sub += 1
        return fast(P)

    def _main(P):
        ctx.sub = 0
        return main(P)

    pr.get_proposal_slow, pr.get_proposal_fast, pr.get_proposal = _slow, _fast, _main
    return ctx


def run_reference(info, n_proposals, seed, chain_id):
    from cobaya.model import get_model
    from cobaya.sampler import get_sampler

    model = get_model(info)
    sampler = get_sampler(info["sampler"], model)
    ctx = install_philox(sampler, seed, chain_id)
    x0 = sampler.current_point.values.copy()
    sampler.n_steps_raw = 0
    for t in range(n_proposals):  # the body of MCMC.run (mcmc.py:470-472)
        ctx.t = t
        sampler.get_new_sample()
        sampler.n_steps_raw += 1
    sampler.collection._cache_dump()
    rows = sampler.collection.data.to_numpy(dtype=np.float64)
    return model, sampler, x0, rows


# ---------------------------------------------------------------------------------
def _load_reference_test_constants():
    """Read ``fixed_info`` (tests/common_sampler.py:24-50) and the deliberately bad
    initial covmat (tests/test_mcmc.py:30-36) from the reference's own test files at
    generation time (constants are not copied into this repo's sources)."""
    import ast

    ns = {"np": np}
    tree = ast.parse(open("/root/reference/tests/common_sampler.py").read())
    for node in tree.body:
        tgt = getattr(node, "target", None) or (getattr(node, "targets", [None])[0])
        if getattr(tgt, "id", "") == "fixed_info":
            value = node.value
            ns["fixed_info"] = eval(compile(ast.Expression(value), "x", "eval"), ns)
    tree = ast.parse(open("/root/reference/tests/test_mcmc.py").read())
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == "test_mcmc":
            for node in fn.body:
                if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "cov":
                    ns["cov0"] = eval(compile(ast.Expression(node.value), "x", "eval"), ns)
    return ns["fixed_info"], ns["cov0"]


def info_g1():
    """configs[0]: the reference's own 3-D test problem (tests/test_mcmc.py:22-82),
    derived parameters on, bad initial covmat."""
    import copy

    fixed_info, cov0 = _load_reference_test_constants()
    info = copy.deepcopy(fixed_info)
    gm = info["likelihood"]["gaussian_mixture"]
    mean, cov = np.array(gm["means"][0]), np.array(gm["covs"][0])
    info["sampler"] = {"mcmc": {"covmat": cov0, "covmat_params": ["a__0", "a__1", "a__2"],
                                "learn_proposal": False, "measure_speeds": False,
                                "burn_in": 0, "seed": 3}}
    return info, mean, cov, cov0


def info_g2():
    """5-D, 2 blocks with oversampling + thinning, 2-mode mixture with weights and
    derived parameters, one normal prior, one periodic parameter, burn-in, T=2."""
    rng = np.random.default_rng(11)
    D = 5
    means = np.array([[0.1, -0.2, 0.3, 0.0, 0.2], [0.35, 0.1, 0.1, 0.25, -0.1]])
    covs = []
    for k in range(2):
        A = rng.standard_normal((D, 2 * D))
        C = A @ A.T / (2 * D)
        s = 0.05 * (1 + rng.uniform(0, 1, D))
        d = np.sqrt(np.diag(C))
        covs.append((C / d[:, None] / d[None, :]) * s[:, None] * s[None, :])
    names = [f"p{i}" for i in range(D)]
    params = {n: {"prior": {"min": -1, "max": 1}, "ref": 0.1} for n in names}
    params["p1"] = {"prior": {"dist": "norm", "loc": 0.0, "scale": 0.5}, "ref": 0.0}
    params["p3"] = {"prior": {"min": -0.5, "max": 0.5}, "ref": 0.1, "periodic": True}
    derived = [f"d{i}" for i in range(2 * D)]
    for dn in derived:
        params[dn] = None
    S0 = np.diag([0.05, 0.06, 0.04, 0.07, 0.05]) ** 2
    S0[0, 1] = S0[1, 0] = 0.3 * 0.05 * 0.06
    S0[2, 4] = S0[4, 2] = -0.4 * 0.04 * 0.05
    return {
        "likelihood": {"gaussian_mixture": {
            "means": means.tolist(), "covs": [c.tolist() for c in covs],
            "weights": [0.3, 0.7], "input_params": names, "output_params": derived,
            "derived": True}},
        "params": params,
        "sampler": {"mcmc": {"covmat": S0, "covmat_params": names,
                             "blocking": [[1, ["p4", "p0", "p2"]], [3, ["p1", "p3"]]],
                             "oversample_thin": True, "temperature": 2,
                             "learn_proposal": False, "measure_speeds": False,
                             "burn_in": 5, "seed": 4}},
    }, means, np.array(covs), S0


def info_g3():
    """4-D dragging: slow block [s0, s1] + fast block [f0, f1] (factor 6)."""
    mean = np.array([0.2, 0.0, 0.1, -0.1])
    rng = np.random.default_rng(5)
    A = rng.standard_normal((4, 8))
    C = A @ A.T / 8
    d = np.sqrt(np.diag(C))
    sig = np.array([0.29, 0.4, 0.2, 0.3])
    cov = (C / d[:, None] / d[None, :]) * sig[:, None] * sig[None, :]
    names = ["s0", "s1", "f0", "f1"]
    params = {n: {"prior": {"min": -3, "max": 3}, "ref": float(mean[i])}
              for i, n in enumerate(names)}
    params["f1"]["periodic"] = True
    S0 = np.diag(sig**2) * 0.5
    return {
        "likelihood": {"gaussian_mixture": {"means": [mean.tolist()],
                                            "covs": [cov.tolist()],
                                            "input_params": names, "derived": False}},
        "params": params,
        "sampler": {"mcmc": {"covmat": S0, "covmat_params": names, "drag": True,
                             "blocking": [[1, ["s0", "s1"]], [6, ["f0", "f1"]]],
                             "learn_proposal": False, "measure_speeds": False,
                             "burn_in": 0, "seed": 5}},
    }, mean, cov, S0


def info_g4():
    """3-D with a 1-parameter block (RandProposer1D) + 2-parameter block."""
    info, mean, cov, cov0 = info_g1()
    info["sampler"]["mcmc"].update(
        blocking=[[1, ["a__2"]], [2, ["a__0", "a__1"]]], oversample_thin=False, seed=6)
    return info, mean, cov, cov0


def info_g5():
    """11-D: one parameter per recognised scipy.stats 1-D prior (cobaya/prior.py:520-525,
    tools.py:611-718), two blocks; the likelihood is wide enough that every prior shapes
    the posterior."""
    names = ["u", "n", "tn", "hn", "ex", "be", "ga", "ln", "ca", "la", "lu"]
    priors = {
        "u": {"min": -1, "max": 2},
        "n": {"dist": "norm", "loc": 0.3, "scale": 0.7},
        "tn": {"dist": "truncnorm", "loc": 0.5, "scale": 0.8, "min": -0.4, "max": 1.9},
        "hn": {"dist": "halfnorm", "loc": 0.1, "scale": 1.3},
        "ex": {"dist": "expon", "loc": -0.2, "scale": 0.9},
        "be": {"dist": "beta", "a": 2.5, "b": 1.7, "min": -0.5, "max": 2.5},
        "ga": {"dist": "gamma", "a": 3.2, "loc": 0.0, "scale": 0.4},
        "ln": {"dist": "lognorm", "s": 0.6, "loc": -0.1, "scale": 1.1},
        "ca": {"dist": "cauchy", "loc": 0.2, "scale": 0.5},
        "la": {"dist": "laplace", "loc": 0.4, "scale": 0.6},
        "lu": {"dist": "loguniform", "a": 0.05, "b": 20.0},
    }
    mean = np.array([0.5, 0.3, 0.6, 0.9, 0.5, 1.2, 1.0, 0.9, 0.2, 0.4, 1.0])
    rng = np.random.default_rng(55)
    A = rng.standard_normal((11, 22))
    C = A @ A.T / 22
    d = np.sqrt(np.diag(C))
    cov = 0.35**2 * C / d[:, None] / d[None, :]
    cov = (cov + cov.T) / 2
    info = {
        "params": {p: {"prior": priors[p], "ref": float(m), "proposal": 0.2}
                   for p, m in zip(names, mean)},
        "likelihood": {"gaussian_mixture": {"means": [mean.tolist()], "covs": [cov.tolist()],
                                            "input_params": names, "output_params": []}},
        "sampler": {"mcmc": {"blocking": [[1, names[:5]], [2, names[5:]]],
                             "covmat": cov * 0.5, "covmat_params": names,
                             "learn_proposal": False, "measure_speeds": False,
                             "burn_in": 0, "seed": 8}},
    }
    return info, mean, cov, cov * 0.5


def info_g7():
    """72-D (the streamed kernels' territory, D > 64): 2-mode mixture with weights, two blocks of
    36 parameters with oversampling + thinning, one normal prior, burn-in."""
    rng = np.random.default_rng(17)
    D = 72
    means = np.stack([rng.uniform(-0.05, 0.05, D), rng.uniform(-0.05, 0.05, D)])
    covs = []
    for k in range(2):
        A = rng.standard_normal((D, 2 * D))
        C = A @ A.T / (2 * D)
        s = 0.04 * (1 + rng.uniform(0, 1, D))
        d = np.sqrt(np.diag(C))
        covs.append((C / d[:, None] / d[None, :]) * s[:, None] * s[None, :])
    names = [f"p{i}" for i in range(D)]
    params = {n: {"prior": {"min": -1, "max": 1}, "ref": 0.0} for n in names}
    params["p5"] = {"prior": {"dist": "norm", "loc": 0.0, "scale": 0.4}, "ref": 0.0}
    S0 = np.diag(np.full(D, 0.03**2))
    perm = [int(i) for i in rng.permutation(D)]
    slow, fast = [names[i] for i in perm[:36]], [names[i] for i in perm[36:]]
    return {
        "likelihood": {"gaussian_mixture": {
            "means": means.tolist(), "covs": [c.tolist() for c in covs],
            "weights": [0.6, 0.4], "input_params": names, "derived": False}},
        "params": params,
        "sampler": {"mcmc": {"covmat": S0, "covmat_params": names,
                             "blocking": [[1, slow], [2, fast]],
                             "oversample_thin": True, "learn_proposal": False,
                             "measure_speeds": False, "burn_in": 3, "seed": 6}},
    }, means, np.array(covs), S0


def _prior_shapes(model):
    a, b, loc, scale = [], [], [], []
    for pdf in model.prior.pdf:
        shapes, lc, sc = pdf.dist._parse_args(*pdf.args, **pdf.kwds)
        a.append(float(shapes[0]) if len(shapes) > 0 else 0.0)
        b.append(float(shapes[1]) if len(shapes) > 1 else 0.0)
        loc.append(float(lc))
        scale.append(float(sc))
    return np.array(a), np.array(b), np.array(loc), np.array(scale)


def prior_known_answers(model, mean, n=400, seed=77):
    """Known answers of Model.logposterior (model.py:579-678) with the scipy priors:
    points scattered around the likelihood mean, some outside the supports."""
    rng = np.random.default_rng(seed)
    X = mean + 0.8 * rng.standard_normal((n, len(mean)))
    X[::7] = mean + 0.05 * rng.standard_normal((len(X[::7]), len(mean)))
    lp, ll = np.empty(n), np.empty(n)
    for i, x in enumerate(X):
        r = model.logposterior(x)
        lp[i] = r.logpriors[0]
        ll[i] = r.loglikes[0] if len(r.loglikes) else np.nan
    return X, lp, ll


def dump_case(name, info_fn, n_proposals, seed, chain_ids):
    info, means, covs, S0 = info_fn()
    out = {}
    for cid in chain_ids:
        model, sampler, x0, rows = run_reference(info, n_proposals, seed, cid)
        out[f"x0_{cid}"] = x0
        out[f"rows_{cid}"] = rows
        out[f"final_x_{cid}"] = sampler.current_point.values.copy()
        out[f"final_weight_{cid}"] = sampler.current_point.weight
        out[f"final_logpost_{cid}"] = sampler.current_point.logpost
    pr = sampler.proposer
    out.update(
        columns=np.array(list(sampler.collection.columns)),
        sampled=np.array(list(model.parameterization.sampled_params())),
        means=np.atleast_2d(means), covs=np.atleast_3d(np.asarray(covs).T).T
        if np.asarray(covs).ndim == 2 else np.asarray(covs),
        S0=np.asarray(S0), proposal_cov=pr.get_covariance(),
        i_of_j=np.asarray(pr.i_of_j), j_start=np.asarray(pr.j_start),
        oversampling=np.asarray(pr.oversampling_factors),
        block_sizes=np.array([bp.n for bp in pr.proposer]),
        output_thin=sampler.current_point.output_thin,
        cycle_length=sampler.cycle_length,
        burn_in=sampler.burn_in.value, max_tries=sampler.max_tries.value,
        temperature=sampler.temperature, proposal_scale=pr.get_scale(),
        drag=int(bool(sampler.drag)),
        drag_interp_steps=int(getattr(sampler, "drag_interp_steps", 0) or 0),
        i_last_slow_block=int(pr.i_last_slow_block),
        n_proposals=n_proposals, seed=seed, chain_ids=np.array(chain_ids),
        lower=model.prior._lower_limits, upper=model.prior._upper_limits,
        periodic=np.array([i in model.prior._periodic_bounds
                           for i in range(model.prior.d())]),
        prior_dist=np.array([pdf.dist.name for pdf in model.prior.pdf]),
        prior_loc=_prior_shapes(model)[2], prior_scale=_prior_shapes(model)[3],
        prior_a=_prior_shapes(model)[0], prior_b=_prior_shapes(model)[1],
        like_weights=np.atleast_1d(np.asarray(
            model.likelihood["gaussian_mixture"].weights, dtype=np.float64)),
        like_derived=int(bool(model.likelihood["gaussian_mixture"].derived)),
        like_input_params=np.array(list(
            model.likelihood["gaussian_mixture"].input_params)),
    )
    for b, T in enumerate(pr.transform):
        out[f"transform_{b}"] = T
    if name.startswith("g5"):
        X, lp, ll = prior_known_answers(model, np.atleast_2d(means)[0])
        out.update(kat_x=X, kat_logprior=lp, kat_loglike=ll)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()
                 if k.startswith("rows_")})


def info_g6():
    """4-D, two more of the reference's internal likelihoods: ``gaussian`` with
    ``normalized: False`` (likelihoods/gaussian/gaussian.py) and ``one``
    (likelihoods/one/one.py), a normal and a uniform prior."""
    rng = np.random.default_rng(66)
    A = rng.standard_normal((4, 8))
    cov = 0.1**2 * (A @ A.T / 8)
    cov = (cov + cov.T) / 2
    mean = np.array([0.1, -0.2, 0.05, 0.3])
    names = ["q0", "q1", "q2", "q3"]
    info = {
        "params": {p: {"prior": {"min": -1, "max": 1}, "ref": float(m), "proposal": 0.05}
                   for p, m in zip(names, mean)},
        "likelihood": {"gaussian": {"mean": mean.tolist(), "cov": cov.tolist(),
                                    "normalized": False, "input_params": names},
                       "one": None},
        "sampler": {"mcmc": {"covmat": cov, "covmat_params": names, "learn_proposal": False,
                             "measure_speeds": False, "burn_in": 0, "seed": 9}},
    }
    info["params"]["q2"]["prior"] = {"dist": "norm", "loc": 0.0, "scale": 0.5}
    return info, mean, cov


def dump_g6(n_proposals=800, seed=106, cid=4):
    info, mean, cov = info_g6()
    out = {}
    for normalized in (False, True):
        info["likelihood"]["gaussian"]["normalized"] = normalized
        model, sampler, x0, rows = run_reference(info, n_proposals, seed, cid)
        tag = "norm" if normalized else "raw"
        out[f"rows_{tag}"] = rows
        out[f"final_x_{tag}"] = sampler.current_point.values.copy()
        out[f"final_weight_{tag}"] = sampler.current_point.weight
        out["x0"] = x0
    out.update(columns=np.array(list(sampler.collection.columns)),
               sampled=np.array(list(model.parameterization.sampled_params())),
               likes=np.array(list(model.likelihood)),
               mean=mean, cov=cov, proposal_cov=sampler.proposer.get_covariance(),
               n_proposals=n_proposals, seed=seed, chain_id=cid,
               max_tries=sampler.max_tries.value)
    np.savez_compressed(os.path.join(GOLDEN, "g6_gaussian_one.npz"), **out)
    print("g6_gaussian_one", out["rows_raw"].shape, out["rows_norm"].shape)


def dump_g8(n_proposals=1200, seed=108, cids=(0, 6)):
    """External likelihood function (LikelihoodExternalFunction, likelihood.py:150-255): the
    unmodified reference evaluates the PYTHON callable of tests/ext_functions.py; the engine
    evaluates its CUDA twin through the device-functor route."""
    from tests import ext_functions

    info, S0 = ext_functions.info_g8()
    out = {}
    for cid in cids:
        model, sampler, x0, rows = run_reference(info, n_proposals, seed, cid)
        out[f"x0_{cid}"] = x0
        out[f"rows_{cid}"] = rows
        out[f"final_x_{cid}"] = sampler.current_point.values.copy()
        out[f"final_weight_{cid}"] = sampler.current_point.weight
    pr = sampler.proposer
    rng = np.random.default_rng(88)
    pts = np.column_stack([rng.uniform(-2.3, 2.3, 40), rng.uniform(-0.8, 2.5, 40),
                           rng.normal(0, 1, 40)])
    kat = []
    for p_ in pts:
        r = model.logposterior(p_)
        kat.append([r.logpost, r.logprior] + (list(r.loglikes) if len(r.loglikes) else [np.nan] * 2))
    out.update(columns=np.array(list(sampler.collection.columns)),
               sampled=np.array(list(model.parameterization.sampled_params())),
               likes=np.array(list(model.likelihood)), S0=S0,
               proposal_cov=pr.get_covariance(), i_of_j=np.asarray(pr.i_of_j),
               block_sizes=np.array([bp.n for bp in pr.proposer]),
               oversampling=np.asarray(pr.oversampling_factors),
               output_thin=sampler.current_point.output_thin,
               n_proposals=n_proposals, seed=seed, chain_ids=np.array(cids),
               max_tries=sampler.max_tries.value, kat_x=pts, kat=np.array(kat))
    np.savez_compressed(os.path.join(GOLDEN, "g8_external.npz"), **out)
    print("g8_external", {k: v.shape for k, v in out.items() if k.startswith("rows_")},
          "blocks", out["block_sizes"], out["oversampling"], "thin", out["output_thin"])


def dump_g9(n_proposals=1000, seed=109, cids=(3,)):
    """External PRIOR (prior.py:537-577,765-772): the reference evaluates the Python callable of
    tests/ext_functions.py; rows carry its own minuslogprior__ring column."""
    from tests import ext_functions

    info, cov = ext_functions.info_g9()
    out = {}
    for cid in cids:
        model, sampler, x0, rows = run_reference(info, n_proposals, seed, cid)
        out[f"x0_{cid}"] = x0
        out[f"rows_{cid}"] = rows
        out[f"final_x_{cid}"] = sampler.current_point.values.copy()
        out[f"final_weight_{cid}"] = sampler.current_point.weight
    pr = sampler.proposer
    rng = np.random.default_rng(99)
    pts = np.column_stack([rng.uniform(-1.7, 1.7, 40), rng.uniform(-1.4, 1.4, 40),
                           rng.normal(0, 1, 40)])
    kat = []
    for p_ in pts:
        r = model.logposterior(p_)
        kat.append([r.logpost] + list(r.logpriors) + (list(r.loglikes) if len(r.loglikes)
                                                      else [np.nan]))
    out.update(columns=np.array(list(sampler.collection.columns)),
               sampled=np.array(list(model.parameterization.sampled_params())),
               cov=cov, proposal_cov=pr.get_covariance(),
               i_of_j=np.asarray(pr.i_of_j), block_sizes=np.array([bp.n for bp in pr.proposer]),
               oversampling=np.asarray(pr.oversampling_factors),
               output_thin=sampler.current_point.output_thin,
               n_proposals=n_proposals, seed=seed, chain_ids=np.array(cids),
               max_tries=sampler.max_tries.value, kat_x=pts, kat=np.array(kat))
    np.savez_compressed(os.path.join(GOLDEN, "g9_external_prior.npz"), **out)
    print("g9_external_prior", {k: v.shape for k, v in out.items() if k.startswith("rows_")},
          list(out["columns"]))


def dump_units():
    """Known answers from the reference's own functions on fixed inputs."""
    from cobaya.functions import _rvs, inverse_cholesky
    from cobaya.model import get_model
    from cobaya.samplers.mcmc.proposal import BlockedProposer

    out = {}
    # KAT-SO(N): numba _rvs on normals from the engine's BASIS stream
    for n in (2, 3, 7, 16, 64):
        xx = orc.basis_normals(n, 1234, 17, 1, 5)
        H = np.eye(n)
        _rvs(np.int64(n), xx.copy(), H)
        out[f"son_xx_{n}"] = xx
        out[f"son_R_{n}"] = H
    # KAT1 model.logposterior
    info, mean, cov, S0 = info_g1()
    model = get_model(info)
    pts = [mean + np.array([0.01, -0.02, 0.03]), np.array([1.5, 0.0, 0.0]),
           np.array([0.2, 0.3, 0.9]), mean, np.array([-1.0, 1.0, 0.0])]
    lps = []
    for p in pts:
        r = model.logposterior(p)
        lps.append([r.logpost, r.logprior, r.loglike] + list(r.derived if len(r.derived)
                                                              else [np.nan] * 3))
    out["kat1_points"] = np.array(pts)
    out["kat1_results"] = np.array(lps)
    out["kat1_mean"], out["kat1_cov"] = mean, cov
    # KAT2/3 set_covariance
    S = S0
    bp = BlockedProposer([[0, 1, 2]], np.random.default_rng(0), oversampling_factors=[1])
    bp.set_covariance(S)
    out["kat2_cov"], out["kat2_T0"] = S, bp.transform[0]
    S3 = cov
    bp = BlockedProposer([[2], [0, 1]], np.random.default_rng(0),
                         oversampling_factors=[1, 3])
    bp.set_covariance(S3)
    out["kat3_cov"] = S3
    out["kat3_T0"], out["kat3_T1"] = bp.transform[0], bp.transform[1]
    out["kat3_i_of_j"] = bp.i_of_j
    out["kat3_multiset"] = np.asarray(bp.block_cycler.sorted_indices)
    # mixture with weights (logsumexp path)
    rng = np.random.default_rng(3)
    info2, means2, covs2, _ = info_g2()
    model2 = get_model(info2)
    pts2 = rng.uniform(-0.4, 0.4, (6, 5))
    res2 = []
    for p in pts2:
        r = model2.logposterior(p)
        res2.append([r.logpost, r.logprior, r.loglike] + list(r.derived))
    out["kat5_points"], out["kat5_results"] = pts2, np.array(res2)
    out["kat5_linv"] = np.array([inverse_cholesky(c) for c in covs2])
    np.savez_compressed(os.path.join(GOLDEN, "units.npz"), **out)
    print("units", sorted(out))


def dump_checkpoint():
    """The reference's own single-chain checkpoint (mcmc.py:795-889,1009-1030) on a
    chain it produced: R-1 and the learned proposal covariance."""
    info, mean, cov, S0 = info_g1()
    info["sampler"]["mcmc"].update(learn_proposal=True, Rminus1_stop=1e-9,
                                   learn_proposal_Rminus1_max=1e9, max_samples=10**9)
    model, sampler, x0, rows = run_reference(info, 6000, 9, 2)
    n = len(rows)
    sampler.check_convergence_and_learn_proposal()
    out = dict(rows=rows, x0=x0, Rminus1=sampler.Rminus1_last,
               learned_cov=sampler.proposer.get_covariance(),
               split=sampler.Rminus1_single_split, n=n,
               acceptance=sampler.progress["acceptance_rate"].iloc[-1],
               mean_full=sampler.collection.mean(first=n // 2),
               cov_full=sampler.collection.cov(first=n // 2))
    # synthetic multi-chain KAT4 through the same arithmetic
    np.savez_compressed(os.path.join(GOLDEN, "checkpoint.npz"), **out)
    print("checkpoint R-1 =", sampler.Rminus1_last, "rows", n)


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    import logging

    logging.disable(logging.WARNING)
    only = set(sys.argv[1:])  # optional: names of the fixtures to (re)generate

    def want(name):
        return not only or name in only

    if want("units"):
        dump_units()
    for name, fn, n, seed, cids in [
        ("g1_gauss3d", info_g1, 1500, 101, [0, 7]),
        ("g2_blocks_mixture", info_g2, 1500, 102, [3]),
        ("g3_dragging", info_g3, 400, 103, [1]),
        ("g4_block1d", info_g4, 1200, 104, [5]),
        ("g5_scipy_priors", info_g5, 1500, 105, [2]),
        ("g7_stream72", info_g7, 450, 107, [2, 9]),
    ]:
        if want(name):
            dump_case(name, fn, n, seed=seed, chain_ids=cids)
    if want("g6_gaussian_one"):
        dump_g6()
    if want("g8_external"):
        dump_g8()
    if want("g9_external_prior"):
        dump_g9()
    if want("checkpoint"):
        dump_checkpoint()
