"""Shared helpers for the test-suite (no reference import at run time)."""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from cobaya_b200.flatmodel import FlatModel, LikeSpec  # noqa: E402


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def flat_from_golden(g) -> FlatModel:
    """Rebuild the FlatModel of a golden case from the raw reference attributes stored
    in the fixture (prior dist names/bounds, mixture means/covs/weights, blocking)."""
    names = [str(s) for s in g["sampled"]]
    D = len(names)
    kind = np.array([0 if str(d) == "uniform" else 1 for d in g["prior_dist"]], np.int32)
    assert all(str(d) in ("uniform", "norm") for d in g["prior_dist"])
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
        loc=g["prior_loc"], pscale=g["prior_scale"], periodic=g["periodic"].astype(np.int32),
        likes=[lk], blocks=blocks, oversampling=[int(o) for o in g["oversampling"]],
        drag=bool(g["drag"]), i_last_slow_block=int(g["i_last_slow_block"]),
        drag_interp_steps=int(g["drag_interp_steps"]),
        proposal_cov=np.asarray(g["proposal_cov"]), proposal_scale=float(g["proposal_scale"]),
        temperature=float(g["temperature"]), max_tries=int(g["max_tries"]),
        output_thin=int(g["output_thin"]),
    )
