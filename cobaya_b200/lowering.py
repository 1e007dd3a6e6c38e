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

from .flatmodel import (PRIOR_NORMAL, PRIOR_UNIFORM, FlatModel, FlatModelError, LikeSpec)


class UnsupportedModelError(FlatModelError):
    pass


def lower_prior(prior):
    """cobaya/prior.py:459-545 -> (kind, lower, upper, loc, scale, periodic)."""
    D = prior.d()
    if len(getattr(prior, "external", {}) or {}):
        raise UnsupportedModelError(
            "External priors are not supported by the B200 ensemble engine "
            "(SURVEY.md section 8f row 3)."
        )
    kind = np.zeros(D, np.int32)
    loc, scale = np.zeros(D), np.ones(D)
    lower = np.array(prior._lower_limits, dtype=np.float64)
    upper = np.array(prior._upper_limits, dtype=np.float64)
    for i, pdf in enumerate(prior.pdf):
        name = pdf.dist.name
        if name == "uniform":
            kind[i] = PRIOR_UNIFORM
        elif name == "norm":
            kind[i] = PRIOR_NORMAL
            loc[i] = pdf.kwds.get("loc", 0.0)
            scale[i] = pdf.kwds.get("scale", 1.0)
        else:
            raise UnsupportedModelError(
                f"1-D prior '{name}' of parameter '{prior.params[i]}' is not in the "
                "engine's recognised set (uniform, norm)."
            )
    periodic = np.zeros(D, np.int32)
    periodic[list(prior._periodic_bounds)] = 1
    return kind, lower, upper, loc, scale, periodic


def lower_likelihoods(model, sampled):
    likes = []
    for name, like in model.likelihood.items():
        cls = type(like).__name__
        if cls == "GaussianMixture":
            idx = [sampled.index(p) for p in like.input_params]
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
        elif cls == "Rosenbrock" and hasattr(like, "b200_scale"):
            idx = [sampled.index(p) for p in like.input_params]
            likes.append(LikeSpec.rosenbrock(idx, scale=like.b200_scale, name=name))
        else:
            raise UnsupportedModelError(
                f"Likelihood '{name}' ({cls}) cannot be evaluated on the device: the "
                "engine recognises gaussian_mixture (and the built-in Rosenbrock). "
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
    kind, lower, upper, loc, scale, periodic = lower_prior(model.prior)
    likes = lower_likelihoods(model, sampled)
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
        periodic=periodic, likes=likes, blocks=blocks, oversampling=oversampling,
        drag=drag, i_last_slow_block=i_last_slow_block,
        drag_interp_steps=drag_interp_steps, proposal_cov=proposal_cov,
        proposal_scale=proposal_scale, temperature=temperature, max_tries=int(max_tries),
        output_thin=output_thin,
    )
