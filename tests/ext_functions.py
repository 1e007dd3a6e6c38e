"""
External likelihood functions used by the tests and by oracle/make_golden.py (g8): a Python
callable for the reference's ``LikelihoodExternalFunction`` and its CUDA twin for the engine.
"""

import numpy as np

from cobaya_b200.functor import device_function

BANANA_CUDA = r'''
// twisted Gaussian ("banana") in (a, b) with an extra smooth term
extern "C" __device__ double banana(const double *p, int n) {
    const double a = p[0], b = p[1];
    const double u = b - 2.0 * a * a;
    return -0.5 * (a * a / 0.09 + u * u / 0.01) - 0.1 * log1p(exp(-3.0 * a));
}
'''


@device_function(BANANA_CUDA)
def banana(a, b):
    u = b - 2.0 * a * a
    return -0.5 * (a * a / 0.09 + u * u / 0.01) - 0.1 * np.log1p(np.exp(-3.0 * a))


def info_g8():
    """3-D: external function of (a, b) + the reference's 1-D ``gaussian`` on c; two blocks."""
    names = ["a", "b", "c"]
    S0 = np.diag([0.2, 0.1, 0.15]) ** 2
    return {
        "params": {"a": {"prior": {"min": -2, "max": 2}, "ref": 0.1, "proposal": 0.2},
                   "b": {"prior": {"min": -1, "max": 3}, "ref": 0.05, "proposal": 0.1},
                   "c": {"prior": {"dist": "norm", "loc": 0.0, "scale": 1.0}, "ref": 0.2,
                         "proposal": 0.15}},
        "likelihood": {"banana": {"external": banana},
                       "gaussian": {"mean": [0.1], "cov": [[0.04]], "input_params": ["c"],
                                    "normalized": True}},
        "sampler": {"mcmc": {"covmat": S0, "covmat_params": names,
                             "blocking": [[1, ["a", "b"]], [2, ["c"]]],
                             "learn_proposal": False, "measure_speeds": False,
                             "burn_in": 0, "seed": 8}},
    }, S0
