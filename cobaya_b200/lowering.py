"""
Lowering of a live reference ``Model`` (+ the blocking/covmat decided by the
sampler front end) to the engine's :class:`~cobaya_b200.flatmodel.FlatModel`.

Reads the model only through the public objects listed in SURVEY.md section 8b:
``model.prior`` (``_bounds``/limits, ``pdf[i].dist.name/.kwds``, ``_periodic_bounds``,
``external``), ``model.likelihood[name]`` (``GaussianMixture.means/.covs/.weights/
.input_params/.output_params/.derived``) and ``model.parameterization``.

Anything outside the recognised set raises :class:`UnsupportedModelError` -- there is
no CPU fallback (BASELINE.json: north_star).
"""

from __future__ import annotations

import numpy as np

from .flatmodel import (PRIOR_KINDS, PRIOR_UNIFORM, FlatModel, FlatModelError, LikeSpec)


class UnsupportedModelError(FlatModelError):
    pass


def lower_prior(prior):
    """cobaya/prior.py:459-545 -> (kind, lower, upper, loc, scale, periodic, a, b).

    ``uniform`` and ``norm`` are the reference's own fast paths (prior.py:514-533); the
    other recognised scipy.stats distributions (flatmodel.PRIOR_KINDS) are evaluated on
    the device in closed form.  Each lowered parameter is checked against its own
    ``pdf.logpdf`` at a few points of the support before it is accepted."""
    D = prior.d()
    kind = np.zeros(D, np.int32)
    loc, scale = np.zeros(D), np.ones(D)
    pa, pb = np.zeros(D), np.zeros(D)
    lower = np.array(prior._lower_limits, dtype=np.float64)
    upper = np.array(prior._upper_limits, dtype=np.float64)
    for i, pdf in enumerate(prior.pdf):
        name = pdf.dist.name
        if name not in PRIOR_KINDS:
            raise UnsupportedModelError(
                f"1-D prior '{name}' of parameter '{prior.params[i]}' is not in the "
                f"engine's recognised set ({', '.join(sorted(PRIOR_KINDS))})."
            )
        kind[i] = PRIOR_KINDS[name]
        if kind[i] == PRIOR_UNIFORM:
            continue
        # rv_frozen keeps shapes in .args and/or .kwds; let scipy sort them out
        shapes, lc, sc = pdf.dist._parse_args(*pdf.args, **pdf.kwds)
        loc[i], scale[i] = float(lc), float(sc)
        if len(shapes) > 0:
            pa[i] = float(shapes[0])
        if len(shapes) > 1:
            pb[i] = float(shapes[1])
        if len(shapes) > 2:
            raise UnsupportedModelError(f"1-D prior '{name}': too many shape parameters")
    periodic = np.zeros(D, np.int32)
    periodic[list(prior._periodic_bounds)] = 1
    _check_lowered_priors(prior, kind, lower, upper, loc, scale, pa, pb)
    return kind, lower, upper, loc, scale, periodic, pa, pb


def _shape_host(kind, z, a, b):
    """Host mirror of csrc/common.cuh ``prior1d_shape`` (used only for the self-check)."""
    if kind in (1, 2, 3):
        return -z * z / 2
    if kind == 4:
        return -z
    if kind == 5:
        return (a - 1) * np.log(z) + (b - 1) * np.log1p(-z)
    if kind == 6:
        return (a - 1) * np.log(z) - z
    if kind == 7:
        return -np.log(z) - np.log(z) ** 2 / (2 * a * a)
    if kind == 8:
        return -np.log1p(z * z)
    if kind == 9:
        return -abs(z)
    if kind == 10:
        return -np.log(z)
    raise ValueError(kind)


def _check_lowered_priors(prior, kind, lower, upper, loc, scale, pa, pb):
    from .flatmodel import prior_log_norm

    for i, pdf in enumerate(prior.pdf):
        if kind[i] < 2:
            continue
        lo, hi = pdf.ppf(0.2), pdf.ppf(0.8)
        cn = prior_log_norm(int(kind[i]), scale[i], pa[i], pb[i])
        for x in (lo, 0.5 * (lo + hi), hi):
            want = float(pdf.logpdf(x))
            got = cn + _shape_host(int(kind[i]), (x - loc[i]) / scale[i], pa[i], pb[i])
            if not np.isclose(got, want, rtol=1e-10, atol=1e-10):
                raise UnsupportedModelError(
                    f"1-D prior '{pdf.dist.name}' of parameter '{prior.params[i]}' could not "
                    f"be lowered consistently (engine {got!r} vs scipy {want!r} at x={x!r}).")


def _reference_class(module: str, name: str):
    """The reference's own likelihood class, or None when that module is not importable."""
    try:
        import importlib

        return getattr(importlib.import_module(module), name)
    except Exception:
        return None


def _indices(like, name, sampled):
    """Positions of a likelihood's input parameters in the sampled vector; an input that is
    fixed or derived from others is outside what the device evaluates."""
    missing = [p for p in like.input_params if p not in sampled]
    if missing:
        raise UnsupportedModelError(
            f"Likelihood '{name}': input parameter(s) {missing} are not sampled (fixed or "
            "dynamically defined inputs are not supported by the B200 ensemble engine).")
    return [sampled.index(p) for p in like.input_params]


def lower_external_priors(prior, sampled, sources=None):
    """External priors (prior.py:537-577): Python callables under ``prior:``.  Those carrying
    their CUDA twin (cobaya_b200.functor.device_function) become device functors; their
    arguments are the function's parameters in signature order (``ExternalPrior.params``).
    ``sources``: the ``prior`` block of the input (lambda strings are translated)."""
    from .functor import DeviceFunctionError, twin_of

    out = []
    for name, ext in (getattr(prior, "external", {}) or {}).items():
        fn = ext.logp
        try:
            source, entry = twin_of(fn, (sources or {}).get(name), name)
        except DeviceFunctionError as e:
            raise UnsupportedModelError(
                f"External prior '{name}' cannot be evaluated on the device: {e}. Attach its "
                "CUDA source with cobaya_b200.functor.device_function. "
                "No CPU fallback is provided.") from e
        missing = [p for p in ext.params if p not in sampled]
        if missing:
            raise UnsupportedModelError(
                f"External prior '{name}': parameter(s) {missing} are not sampled.")
        out.append(LikeSpec.external([sampled.index(p) for p in ext.params], source, entry,
                                     name=name))
    return out


def lower_likelihoods(model, sampled):
    """Recognised by CLASS (isinstance against the reference's classes), never by name: a
    user class that happens to be called ``GaussianMixture`` is not lowered as the built-in."""
    gm_cls = _reference_class("cobaya.likelihoods.gaussian_mixture", "GaussianMixture")
    g_cls = _reference_class("cobaya.likelihoods.gaussian", "Gaussian")
    one_cls = _reference_class("cobaya.likelihoods.one", "one")
    ext_cls = _reference_class("cobaya.likelihood", "LikelihoodExternalFunction")
    if len(getattr(model, "theory", {}) or {}):
        raise UnsupportedModelError(
            "Theory components are not supported by the B200 ensemble engine: "
            f"{list(model.theory)}. No CPU fallback is provided.")
    likes = []
    for name, like in model.likelihood.items():
        cls = type(like).__name__
        if gm_cls is not None and isinstance(like, gm_cls):
            idx = _indices(like, name, sampled)
            weights = like.weights
            if np.isscalar(weights):
                weights = None
            likes.append(
                LikeSpec.gaussian_mixture(
                    idx, np.asarray(like.means), np.asarray(like.covs), weights,
                    derived=bool(like.derived), name=name,
                    derived_names=list(like.output_params) if like.derived else [],
                )
            )
        elif g_cls is not None and isinstance(like, g_cls):
            idx = _indices(like, name, sampled)
            likes.append(LikeSpec.gaussian(idx, np.asarray(like.mean), np.asarray(like.cov),
                                           normalized=bool(getattr(like, "normalized", True)),
                                           name=name))
        elif one_cls is not None and isinstance(like, one_cls) and not getattr(like, "noise", None):
            likes.append(LikeSpec.constant(0.0, name=name))
        elif ext_cls is not None and isinstance(like, ext_cls):
            # an external function: its CUDA twin (cobaya_b200.functor.device_function) or the
            # automatic translation of its lambda string (functor.cuda_from_lambda)
            from inspect import getfullargspec

            from .functor import DeviceFunctionError, twin_of

            if like.output_params or like.get_requirements():
                raise UnsupportedModelError(
                    f"External likelihood '{name}': derived outputs and requirements are not "
                    "supported on the device.")
            fn = like.external_function
            try:
                source, entry = twin_of(fn, getattr(like, "external", None), name)
            except DeviceFunctionError as e:
                raise UnsupportedModelError(
                    f"External likelihood '{name}' cannot be evaluated on the device: {e}. "
                    "Attach its CUDA source with cobaya_b200.functor.device_function. "
                    "No CPU fallback is provided.") from e
            # p[] follows the function's signature
            order = [a for a in getfullargspec(fn).args if a in like.input_params]
            if sorted(order) != sorted(like.input_params):
                raise UnsupportedModelError(
                    f"External likelihood '{name}': inputs {list(like.input_params)} do not "
                    f"match the function's named arguments {order}.")
            _indices(like, name, sampled)
            likes.append(LikeSpec.external([sampled.index(p_) for p_ in order], source, entry,
                                           name=name))
        elif hasattr(like, "b200_scale") and cls == "Rosenbrock":  # the engine's own built-in
            idx = _indices(like, name, sampled)
            likes.append(LikeSpec.rosenbrock(idx, scale=like.b200_scale, name=name))
        else:
            raise UnsupportedModelError(
                f"Likelihood '{name}' ({cls}) cannot be evaluated on the device: the "
                "engine recognises gaussian_mixture, gaussian, one (without noise), the "
                "built-in Rosenbrock and external functions that carry their CUDA source "
                "(cobaya_b200.functor.device_function). "
                "No CPU fallback is provided."
            )
    return likes


def lower_model(model, sampler=None, *, blocks=None, oversampling=None, drag=False,
                i_last_slow_block=None, drag_interp_steps=0, proposal_cov=None,
                proposal_scale=2.4, temperature=1.0, max_tries=None, output_thin=1):
    """Build the FlatModel.  If ``sampler`` (a reference-style MCMC object whose
    ``initialize`` has run) is given, blocking / covmat / options are read from it."""
    sampled = list(model.parameterization.sampled_params())
    par = model.parameterization
    if any(bool(getattr(par, a, None)) for a in ("_input_funcs",)) and par._input_funcs:
        raise UnsupportedModelError(
            "Dynamically defined (lambda) input parameters are not supported."
        )
    kind, lower, upper, loc, scale, periodic, pa, pb = lower_prior(model.prior)
    likes = lower_likelihoods(model, sampled)
    try:
        prior_sources = dict(model.info().get("prior") or {})
    except Exception:
        prior_sources = {}
    ext_priors = lower_external_priors(model.prior, sampled, prior_sources)
    derived_model = [p for p in par.derived_params()]
    derived_engine = [n for lk in likes for n in lk.derived_names]
    if derived_model != derived_engine:
        raise UnsupportedModelError(
            "Derived parameters other than gaussian_mixture's whitened outputs are not "
            f"supported (model: {derived_model}, engine: {derived_engine})."
        )
    if sampler is not None:
        pr = sampler.proposer
        blocks = [[int(i) for i in pr.i_of_j[js: js + bp.n]]
                  for js, bp in zip(pr.j_start, pr.proposer)]
        oversampling = [int(o) for o in pr.oversampling_factors]
        drag = bool(sampler.drag)
        i_last_slow_block = int(pr.i_last_slow_block)
        drag_interp_steps = int(getattr(sampler, "drag_interp_steps", 0) or 0)
        proposal_cov = pr.get_covariance()
        proposal_scale = float(pr.get_scale())
        temperature = float(sampler.temperature)
        max_tries = sampler.max_tries.value
        output_thin = int(sampler.current_point.output_thin)
    if max_tries is None or not np.isfinite(max_tries):
        max_tries = 2**59
    return FlatModel(
        names=sampled, prior_kind=kind, lower=lower, upper=upper, loc=loc, pscale=scale,
        periodic=periodic, pa=pa, pb=pb, likes=likes, blocks=blocks, oversampling=oversampling,
        drag=drag, i_last_slow_block=i_last_slow_block,
        drag_interp_steps=drag_interp_steps, proposal_cov=proposal_cov,
        proposal_scale=proposal_scale, temperature=temperature, max_tries=int(max_tries),
        output_thin=output_thin, ext_priors=ext_priors,
    )
