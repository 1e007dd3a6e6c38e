"""
The synthetic targets of BASELINE.json ``configs`` as FlatModels:
shared by ``bench.py``, the tools and the parity tests so that "configs[2]" means one thing.

* ``c1`` -- configs[1]: 64-D correlated Gaussian, one block, uniform prior [-1, 1], proposal
  started from the diagonal of the target (so that covariance learning has work to do).
* ``c2`` -- configs[2]: 128-D, two ``gaussian_mixture`` components of 3 modes over disjoint
  parameters (0-31 slow, 32-127 fast) -> two speed blocks, oversampling [1, 3],
  ``oversample_thin`` (output_thin 2).  A single likelihood would be one block
  (SURVEY.md a13), hence two component instances.
* ``c3`` -- configs[3]: 30-D Rosenbrock (builder-defined: the reference has none),
  ``logp = -(1/20) sum_i [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2]``, prior [-5, 5], manual
  slow/fast blocking 10 + 20 and ``drag: True`` (mcmc.py:564-668).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .flatmodel import FlatModel, LikeSpec, synthetic_gaussian_cov


@dataclass
class Problem:
    key: str
    workload: str
    fm: FlatModel
    start: object              # start(n, rank, seed) -> x0 [n, D]
    evals_per_proposal: int    # posterior evaluations per (slow) proposal
    flops_per_proposal: float  # 4 D^2 (+ D^2 per extra mode)
    target_cov: np.ndarray | None = None


def mixture_cov(D, rng, scale=0.02):
    A = rng.standard_normal((D, 2 * D))
    C = A @ A.T / (2 * D)
    d = np.sqrt(np.diag(C))
    s = scale * 10 ** rng.uniform(-0.5, 0.5, D)
    return (C / d[:, None] / d[None, :]) * s[:, None] * s[None, :]


def config1(D: int = 64) -> Problem:
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, bounds=(-1.0, 1.0),
                            proposal_cov=np.diag(np.diag(cov)))

    def start(n, rank, seed=1):
        return np.random.default_rng([seed, rank]).multivariate_normal(np.zeros(D), cov, size=n)

    return Problem("c1", f"{D}-D correlated Gaussian, 8192 chains/GPU, covmat learning on "
                         "(BASELINE configs[1])", fm, start, 1, 4.0 * D * D, cov)


def config2() -> Problem:
    rng = np.random.default_rng(20260925)
    D, n_slow = 128, 32
    a = LikeSpec.gaussian_mixture(np.arange(n_slow),
                                  [np.full(n_slow, 0.03 * k) for k in range(3)],
                                  [mixture_cov(n_slow, rng) for _ in range(3)], name="slow")
    b = LikeSpec.gaussian_mixture(np.arange(n_slow, D),
                                  [np.full(D - n_slow, 0.03 * k) for k in range(3)],
                                  [mixture_cov(D - n_slow, rng) for _ in range(3)], name="fast")
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -1.0), upper=np.full(D, 1.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[a, b],
                   blocks=[list(range(n_slow)), list(range(n_slow, D))], oversampling=[1, 3],
                   proposal_cov=np.diag(np.full(D, 0.01 ** 2)), output_thin=2)

    def start(n, rank, seed=1):
        return np.random.default_rng([seed, rank]).normal(0, 0.01, (n, D))

    return Problem("c2", "128-D gaussian_mixture, 2 components x 3 modes, speed blocks 32+96, "
                         "oversampling [1,3], 8192 chains/GPU (BASELINE configs[2])",
                   fm, start, 1, 4.0 * D * D + 2.0 * (n_slow ** 2 + (D - n_slow) ** 2))


def config3() -> Problem:
    D, n_slow, o_fast = 30, 10, 4
    lk = LikeSpec.rosenbrock(np.arange(D), scale=1.0 / 20.0)
    n_drag = int(np.round(o_fast * (D - n_slow) / n_slow))
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -5.0), upper=np.full(D, 5.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=[list(range(n_slow)), list(range(n_slow, D))],
                   oversampling=[1, o_fast], drag=True, i_last_slow_block=0,
                   drag_interp_steps=n_drag, proposal_cov=np.diag(np.full(D, 0.05 ** 2)))

    def start(n, rank, seed=1):
        return 1.0 + np.random.default_rng([seed, rank]).normal(0, 0.05, (n, D))

    return Problem("c3", f"30-D Rosenbrock, dragging (n_drag = {n_drag}), blocks 10+20, 8192 "
                         "chains/GPU (BASELINE configs[3])", fm, start, 2 * n_drag + 1,
                   (2 * n_drag + 1) * 6.0 * D + 4.0 * D * D)


ROSENBROCK_CUDA = r'''
// -scale * sum_i [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2]  (BASELINE configs[3])
extern "C" __device__ double rosenbrock_ext(const double *p, int n) {
    double acc = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
        const double t1 = p[i + 1] - p[i] * p[i], t2 = 1.0 - p[i];
        acc += 100.0 * t1 * t1 + t2 * t2;
    }
    return -(1.0 / 20.0) * acc;
}
'''


def config3_external() -> Problem:
    """configs[3] literally: the Rosenbrock function as an EXTERNAL likelihood (CUDA source
    compiled with NVRTC at run time), dragging through the split-launch route."""
    p = config3()
    p.fm.likes = [LikeSpec.external(np.arange(p.fm.D), ROSENBROCK_CUDA, "rosenbrock_ext",
                                    name=p.fm.likes[0].name)]
    return Problem("c3x", p.workload.replace("(BASELINE", "as an external CUDA function "
                                                          "(NVRTC), split launches (BASELINE"),
                   p.fm, p.start, p.evals_per_proposal, p.flops_per_proposal)


def get(key: str) -> Problem:
    return {"c1": config1, "c2": config2, "c3": config3, "c3x": config3_external}[key]()
