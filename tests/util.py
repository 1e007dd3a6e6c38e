"""Shared helpers for the test-suite (no reference import at run time)."""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from cobaya_b200.flatmodel import PRIOR_KINDS, FlatModel, LikeSpec  # noqa: E402


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def flat_from_golden(g) -> FlatModel:
    """Rebuild the FlatModel of a golden case from the raw reference attributes stored
    in the fixture (prior dist names/bounds, mixture means/covs/weights, blocking)."""
    names = [str(s) for s in g["sampled"]]
    D = len(names)
    kind = np.array([PRIOR_KINDS[str(d)] for d in g["prior_dist"]], np.int32)
    pa = g["prior_a"] if "prior_a" in g.files else np.zeros(D)
    pb = g["prior_b"] if "prior_b" in g.files else np.zeros(D)
    idx = [names.index(str(p)) for p in g["like_input_params"]]
    w = np.asarray(g["like_weights"])
    means = np.atleast_2d(g["means"])
    covs = np.asarray(g["covs"]).reshape(means.shape[0], D, D)
    weights = None if len(w) == 1 and means.shape[0] != 1 else (w if len(w) > 1 else None)
    cols = [str(c) for c in g["columns"]]
    der_names = cols[2 + D: cols.index("minuslogprior")]
    lk = LikeSpec.gaussian_mixture(idx, means, covs, weights,
                                   derived=bool(g["like_derived"]),
                                   name="gaussian_mixture", derived_names=der_names)
    i_of_j = [int(i) for i in g["i_of_j"]]
    blocks, p = [], 0
    for n in g["block_sizes"]:
        blocks.append(i_of_j[p: p + int(n)])
        p += int(n)
    return FlatModel(
        names=names, prior_kind=kind, lower=g["lower"], upper=g["upper"],
        loc=g["prior_loc"], pscale=g["prior_scale"], pa=pa, pb=pb,
        periodic=g["periodic"].astype(np.int32),
        likes=[lk], blocks=blocks, oversampling=[int(o) for o in g["oversampling"]],
        drag=bool(g["drag"]), i_last_slow_block=int(g["i_last_slow_block"]),
        drag_interp_steps=int(g["drag_interp_steps"]),
        proposal_cov=np.asarray(g["proposal_cov"]), proposal_scale=float(g["proposal_scale"]),
        temperature=float(g["temperature"]), max_tries=int(g["max_tries"]),
        output_thin=int(g["output_thin"]),
    )


def flat_g6(g, normalized: bool) -> FlatModel:
    """The g6 fixture: ``gaussian`` (normalized or not) + ``one`` likelihoods, uniform
    priors on [-1, 1] except a normal prior on q2 (oracle/make_golden.py info_g6)."""
    names = [str(s) for s in g["sampled"]]
    D = len(names)
    kind = np.zeros(D, np.int32); kind[2] = 1
    lower, upper = np.full(D, -1.0), np.full(D, 1.0)
    lower[2], upper[2] = -np.inf, np.inf
    sc = np.ones(D); sc[2] = 0.5
    likes = [LikeSpec.gaussian(np.arange(D), g["mean"], g["cov"], normalized=normalized,
                               name="gaussian"), LikeSpec.constant(0.0, name="one")]
    assert [str(n) for n in g["likes"]] == ["gaussian", "one"]
    return FlatModel(names=names, prior_kind=kind, lower=lower, upper=upper, loc=np.zeros(D),
                     pscale=sc, periodic=np.zeros(D, np.int32), likes=likes,
                     proposal_cov=np.asarray(g["proposal_cov"]), max_tries=int(g["max_tries"]))


def flat_g8(g, source=None):
    from cobaya_b200.flatmodel import FlatModel, LikeSpec

    from tests import ext_functions

    names = [str(s) for s in g["sampled"]]
    assert names == ["a", "b", "c"] and [str(s) for s in g["likes"]] == ["banana", "gaussian"]
    kind = np.array([0, 0, 1], np.int32)
    lower = np.array([-2.0, -1.0, -np.inf])
    upper = np.array([2.0, 3.0, np.inf])
    likes = [LikeSpec.external([0, 1], source or ext_functions.BANANA_CUDA, "banana",
                               name="banana"),
             LikeSpec.gaussian([2], [0.1], [[0.04]], normalized=True, name="gaussian")]
    i_of_j = [int(i) for i in g["i_of_j"]]
    blocks, j = [], 0
    for n in g["block_sizes"]:
        blocks.append(i_of_j[j: j + int(n)])
        j += int(n)
    return FlatModel(names=names, prior_kind=kind, lower=lower, upper=upper, loc=np.zeros(3),
                     pscale=np.ones(3), periodic=np.zeros(3, np.int32), likes=likes,
                     blocks=blocks, oversampling=[int(o) for o in g["oversampling"]],
                     output_thin=int(g["output_thin"]),
                     proposal_cov=np.asarray(g["proposal_cov"]), max_tries=int(g["max_tries"]))
