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


RING_CUDA = r'''
// a soft ring prior on (a, b): log N(sqrt(a^2 + b^2); 0.5, 0.1), not normalised
extern "C" __device__ double ring(const double *p, int n) {
    const double rr = sqrt(p[0] * p[0] + p[1] * p[1]);
    return -0.5 * (rr - 0.5) * (rr - 0.5) / 0.01;
}
'''


@device_function(RING_CUDA)
def ring(a, b):
    rr = np.sqrt(a * a + b * b)
    return -0.5 * (rr - 0.5) ** 2 / 0.01


def info_g9():
    """3-D: the reference's ``gaussian_mixture`` on (a, b, c) with an EXTERNAL PRIOR on (a, b)
    (prior.py:537-577) and a normal 1-D prior on c."""
    names = ["a", "b", "c"]
    cov = np.diag([0.3, 0.3, 0.2]) ** 2
    cov[0, 1] = cov[1, 0] = 0.03
    return {
        "params": {"a": {"prior": {"min": -1.5, "max": 1.5}, "ref": 0.5, "proposal": 0.1},
                   "b": {"prior": {"min": -1.5, "max": 1.5}, "ref": 0.05, "proposal": 0.1},
                   "c": {"prior": {"dist": "norm", "loc": 0.0, "scale": 1.0}, "ref": 0.1,
                         "proposal": 0.1}},
        "prior": {"ring": ring},
        "likelihood": {"gaussian_mixture": {"means": [[0.1, 0.0, 0.05]], "covs": [cov.tolist()],
                                            "input_params": names, "derived": False}},
        "sampler": {"mcmc": {"covmat": np.diag([0.1, 0.1, 0.1]) ** 2, "covmat_params": names,
                             "learn_proposal": False, "measure_speeds": False,
                             "burn_in": 0, "seed": 9}},
    }, cov


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


GAUSS_CUDA = r'''
// the same Gaussian as the built-in mixture: -0.5 (D log 2pi + logdet + |Linv (x - mu)|^2)
__device__ const double LINV[%(D)d][%(D)d] = {%(linv)s};
__device__ const double MU[%(D)d] = {%(mu)s};
extern "C" __device__ double gauss_ext(const double *p, int n) {
    double q = 0.0;
    for (int i = 0; i < %(D)d; ++i) {
        double a = 0.0;
        for (int j = 0; j <= i; ++j) a += LINV[i][j] * (p[j] - MU[j]);
        q += a * a;
    }
    return -0.5 * (%(c0).17g + q);
}
'''


def gaussian_pair(D):
    """The benchmark's correlated Gaussian twice: as the built-in mixture and as an external
    CUDA function (same proposals, log-likelihoods equal up to the summation order)."""
    from cobaya_b200.flatmodel import FlatModel, LikeSpec, synthetic_gaussian_cov

    cov = synthetic_gaussian_cov(D)
    mu = np.linspace(-0.01, 0.01, D)
    builtin = FlatModel.gaussian(mu[None], cov[None], proposal_cov=cov, bounds=(-1.0, 1.0))
    lk = builtin.likes[0]
    src = GAUSS_CUDA % dict(
        D=D,
        linv=", ".join("{" + ", ".join(f"{v:.17g}" for v in row) + "}" for row in lk.linv[0]),
        mu=", ".join(f"{v:.17g}" for v in mu),
        c0=D * np.log(2 * np.pi) + lk.logdet[0])
    ext = FlatModel(names=list(builtin.names), prior_kind=builtin.prior_kind,
                    lower=builtin.lower, upper=builtin.upper, loc=builtin.loc,
                    pscale=builtin.pscale, periodic=builtin.periodic,
                    likes=[LikeSpec.external(np.arange(D), src, "gauss_ext",
                                             name="gaussian_mixture")],
                    proposal_cov=cov)
    return builtin, ext, mu, cov


from cobaya_b200.problems import ROSENBROCK_CUDA  # noqa: E402


def rosenbrock_pair():
    """configs[3] twice: with the built-in Rosenbrock and with the same function as an external
    CUDA function (dragging, 10 slow + 20 fast parameters)."""
    import copy

    from cobaya_b200 import problems
    from cobaya_b200.flatmodel import LikeSpec

    p = problems.config3()
    ext = copy.deepcopy(p.fm)
    ext.likes = [LikeSpec.external(np.arange(ext.D), ROSENBROCK_CUDA, "rosenbrock_ext",
                                   name=p.fm.likes[0].name)]
    return p.fm, ext, p.start
