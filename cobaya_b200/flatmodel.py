"""
Flat (array-only) description of the posterior + proposal that the CUDA engine
consumes.  This is the *lowered* form of what the reference keeps in Python
objects:

* ``Prior``            -> cobaya/prior.py:514-533 (bounds, uniform/normal kinds,
                          ``_uniform_logp``), periodic flags (:500-513)
* ``GaussianMixture``  -> cobaya/likelihoods/gaussian_mixture/gaussian_mixture.py:45-136
                          (means, covs -> inverse Cholesky factors, weights)
* ``BlockedProposer``  -> cobaya/samplers/mcmc/proposal.py:96-260 (blocks,
                          ``i_of_j``, oversampling, transforms)

Host-side only: numpy/LAPACK D x D setup work, done once per run or once per
covariance learn (mcmc.py:1023), exactly where the reference does it.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from itertools import chain as _chain
from typing import Sequence

import numpy as np
from scipy.linalg import lapack as _lapack

LIKE_GAUSSIAN_MIXTURE = 0
LIKE_ROSENBROCK = 1
LIKE_CONSTANT = 2
LIKE_EXTERNAL = 3
MAX_BLOCKS = 16

PRIOR_UNIFORM = 0
PRIOR_NORMAL = 1
# scipy.stats distributions the reference reaches through ``pdf.logpdf``
# (prior.py:520-525); numbering = include/cobaya_b200.h CB2_PRIOR_*
PRIOR_KINDS = {"uniform": 0, "norm": 1, "truncnorm": 2, "halfnorm": 3, "expon": 4, "beta": 5,
               "gamma": 6, "lognorm": 7, "cauchy": 8, "laplace": 9, "loguniform": 10,
               "reciprocal": 10}


def _log_gauss_mass(a: float, b: float) -> float:
    """log(Phi(b) - Phi(a)), evaluated on the side where the difference does not cancel."""
    from scipy.special import log_ndtr, ndtr

    if b <= 0:
        return float(log_ndtr(b) + np.log1p(-np.exp(log_ndtr(a) - log_ndtr(b))))
    if a >= 0:
        return _log_gauss_mass(-b, -a)
    return float(np.log1p(-ndtr(a) - ndtr(-b)))


def prior_log_norm(kind: int, scale: float, a: float, b: float) -> float:
    """Additive constant ``cn`` of ``logpdf(x) = cn + f((x-loc)/scale; a, b)`` for the
    engine's generic 1-D priors (f: csrc/common.cuh ``prior1d_shape``)."""
    from scipy.special import betaln, gammaln

    ls = np.log(scale)
    half_log_2pi = 0.5 * np.log(2 * np.pi)
    if kind == 2:
        if not a < b:
            raise FlatModelError("truncnorm needs a < b")
        return float(-half_log_2pi - _log_gauss_mass(a, b) - ls)
    if kind == 3:
        return float(0.5 * np.log(2 / np.pi) - ls)
    if kind == 4:
        return float(-ls)
    if kind == 5:
        if not (a > 0 and b > 0):
            raise FlatModelError("beta needs a, b > 0")
        return float(-betaln(a, b) - ls)
    if kind == 6:
        if not a > 0:
            raise FlatModelError("gamma needs a > 0")
        return float(-gammaln(a) - ls)
    if kind == 7:
        if not a > 0:
            raise FlatModelError("lognorm needs s > 0")
        return float(-np.log(a) - half_log_2pi - ls)
    if kind == 8:
        return float(-np.log(np.pi) - ls)
    if kind == 9:
        return float(-np.log(2.0) - ls)
    if kind == 10:
        if not 0 < a < b:
            raise FlatModelError("loguniform needs 0 < a < b")
        return float(-np.log(np.log(b) - np.log(a)) - ls)
    return 0.0


class FlatModelError(ValueError):
    """Raised when a model cannot be lowered to the engine's recognised set."""


def inverse_cholesky(cov: np.ndarray) -> np.ndarray:
    """L^-1 with cov = L L^T (mirrors cobaya/functions.py:81-89)."""
    chol = np.linalg.cholesky(np.asarray(cov, dtype=np.float64))
    linv, info = _lapack.dtrtri(chol, lower=True)
    if info != 0:
        raise np.linalg.LinAlgError("dtrtri failed")
    return np.tril(linv)


def cov_to_std_and_corr(cov: np.ndarray):
    """Mirrors cobaya/tools.py:779-788."""
    std = np.sqrt(np.diag(cov))
    inv_std = 1 / std
    corr = inv_std[:, np.newaxis] * cov * inv_std[np.newaxis, :]
    np.fill_diagonal(corr, 1.0)
    return std, corr


def transforms_from_cov(cov: np.ndarray, i_of_j: np.ndarray) -> np.ndarray:
    """
    Full lower-triangular transform ``T = S L'`` in *block-sorted* coordinates
    (proposal.py:250-260, tools.py:761-776).  ``transform[b]`` of the reference
    is the slice ``T[j_b:, j_b:j_b+n_b]``; a block-b proposal is ``T @ v`` with
    ``v`` supported on block b's columns.
    """
    cov = np.asarray(cov, dtype=np.float64)
    if cov.shape != (len(i_of_j), len(i_of_j)):
        raise FlatModelError(
            "The covariance matrix does not have the correct dimension: "
            f"it's {cov.shape[0]}, but it should be {len(i_of_j)}."
        )
    if not (np.allclose(cov.T, cov) and np.all(np.linalg.eigvalsh((cov + cov.T) / 2) > 0)):
        raise FlatModelError(
            "The given covmat is not a positive-definite, symmetric square matrix."
        )
    sorted_cov = cov[np.ix_(i_of_j, i_of_j)]
    std, corr = cov_to_std_and_corr(sorted_cov)
    Lp = np.linalg.cholesky(corr)
    return np.tril(np.diag(std).dot(Lp))


@dataclass
class LikeSpec:
    kind: int
    idx: np.ndarray  # [dim] indices into the sampled vector
    name: str = "like"
    # gaussian mixture
    means: np.ndarray | None = None  # [m, dim]
    covs: np.ndarray | None = None  # [m, dim, dim]
    linv: np.ndarray | None = None  # [m, dim, dim]
    logdet: np.ndarray | None = None  # [m]
    weights: np.ndarray | None = None  # [m]
    derived: bool = False
    derived_names: list = field(default_factory=list)
    # rosenbrock
    scale: float = 1.0
    # external function (device functor): CUDA source and the name of its entry point
    source: str | None = None
    fn_name: str | None = None

    @property
    def dim(self) -> int:
        return len(self.idx)

    @property
    def n_modes(self) -> int:
        return 0 if self.means is None else self.means.shape[0]

    @classmethod
    def gaussian_mixture(cls, idx, means, covs, weights=None, derived=False, name="gm",
                         derived_names=None):
        idx = np.asarray(idx, dtype=np.int32)
        means = np.atleast_2d(np.asarray(means, dtype=np.float64))
        covs = np.asarray(covs, dtype=np.float64)
        if covs.ndim == 2:
            covs = covs[None]
        m, d = means.shape
        if covs.shape != (m, d, d) or d != len(idx):
            raise FlatModelError("gaussian_mixture: inconsistent means/covs/params shapes")
        if weights is None or (np.isscalar(weights) and not weights):
            w = np.full(m, 1.0 / m)  # gaussian_mixture.py:133
        else:
            w = np.asarray(weights, dtype=np.float64).reshape(-1)
            if len(w) != m:
                raise FlatModelError("There must be as many weights as components.")
            if not np.isclose(w.sum(), 1):
                w = w / w.sum()  # gaussian_mixture.py:129-131
        linv = np.stack([inverse_cholesky(c) for c in covs])
        # scipy's multivariate_normal uses the eigen-decomposition log-pdet
        logdet = np.array([np.sum(np.log(np.linalg.eigvalsh(c))) for c in covs])
        return cls(
            kind=LIKE_GAUSSIAN_MIXTURE, idx=idx, name=name, means=means, covs=covs,
            linv=linv, logdet=logdet, weights=w, derived=bool(derived),
            derived_names=list(derived_names or []),
        )

    @classmethod
    def rosenbrock(cls, idx, scale=1.0 / 20.0, name="rosenbrock"):
        return cls(kind=LIKE_ROSENBROCK, idx=np.asarray(idx, dtype=np.int32), name=name,
                   scale=float(scale))

    @classmethod
    def gaussian(cls, idx, mean, cov, normalized=True, name="gaussian"):
        """``gaussian`` (likelihoods/gaussian/gaussian.py:96-112): -chi2/2 + log_norm with
        log_norm = -(k log 2pi + log|cov|)/2, or 0 if not ``normalized`` -- a one-mode
        mixture whose normalisation is fixed through the log-determinant entry."""
        lk = cls.gaussian_mixture(idx, np.atleast_1d(mean), np.atleast_2d(cov), name=name)
        d = lk.dim
        if normalized:
            sign, logdet = np.linalg.slogdet(np.atleast_2d(cov))  # gaussian.py:84
            if sign <= 0:
                raise FlatModelError("The covariance matrix is not positive definite!")
            lk.logdet = np.array([logdet])
        else:
            lk.logdet = np.array([-d * np.log(2 * np.pi)])
        return lk

    @classmethod
    def external(cls, idx, source, fn_name, name="external"):
        """An external likelihood function (likelihood.py:150-255) as a device functor:
        ``source`` defines ``extern "C" __device__ double fn_name(const double *p, int n)``,
        ``p`` holding the input parameters in the order of ``idx``; it returns log L."""
        return cls(kind=LIKE_EXTERNAL, idx=np.asarray(idx, dtype=np.int32), name=name,
                   source=str(source), fn_name=str(fn_name))

    @classmethod
    def constant(cls, value=0.0, name="one"):
        """``one`` (likelihoods/one/one.py:26-28)."""
        return cls(kind=LIKE_CONSTANT, idx=np.zeros(0, np.int32), name=name, scale=float(value))

    @property
    def n_derived(self) -> int:
        return self.dim * self.n_modes if (self.kind == LIKE_GAUSSIAN_MIXTURE and
                                           self.derived) else 0


@dataclass
class FlatModel:
    """Everything the engine (and the oracle) needs, as plain arrays."""

    names: list
    prior_kind: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    loc: np.ndarray
    pscale: np.ndarray
    periodic: np.ndarray
    likes: list
    # scipy shape parameters of the generic 1-D priors (kinds >= 2)
    pa: np.ndarray | None = None
    pb: np.ndarray | None = None
    # blocking: list of blocks (lists of sampler indices), ascending speed
    blocks: list = None
    oversampling: list = None
    drag: bool = False
    i_last_slow_block: int | None = None
    drag_interp_steps: int = 0
    # proposal
    proposal_cov: np.ndarray | None = None
    proposal_scale: float = 2.4
    # options
    temperature: float = 1.0
    max_tries: int = 2**59  # 'no limit'; the sampler front ends set mcmc.yaml's value
    output_thin: int = 1
    # external priors (prior.py:537-577) as device functors: LikeSpec.external entries
    ext_priors: list = field(default_factory=list)

    def __post_init__(self):
        D = len(self.names)
        self.prior_kind = np.asarray(self.prior_kind, dtype=np.int32).reshape(D)
        for a in ("lower", "upper", "loc", "pscale"):
            setattr(self, a, np.asarray(getattr(self, a), dtype=np.float64).reshape(D))
        self.periodic = np.asarray(self.periodic, dtype=np.int32).reshape(D)
        for a in ("pa", "pb"):
            v = getattr(self, a)
            setattr(self, a, np.zeros(D) if v is None else
                    np.asarray(v, dtype=np.float64).reshape(D))
        if np.any((self.prior_kind < 0) | (self.prior_kind > max(PRIOR_KINDS.values()))):
            raise FlatModelError("unknown 1-D prior kind")
        self.prior_log_norm = np.array([
            prior_log_norm(int(k), s, a, b) if k >= 2 else 0.0
            for k, s, a, b in zip(self.prior_kind, self.pscale, self.pa, self.pb)])
        if self.blocks is None:
            self.blocks = [list(range(D))]
        if self.oversampling is None:
            self.oversampling = [1] * len(self.blocks)
        if len(self.blocks) > MAX_BLOCKS:
            raise FlatModelError(f"at most {MAX_BLOCKS} parameter blocks are supported")
        if len(self.oversampling) != len(self.blocks):
            raise FlatModelError(
                "List of oversampling factors has a different length that list of blocks"
            )
        if set(_chain(*self.blocks)) != set(range(D)) or sum(map(len, self.blocks)) != D:
            raise FlatModelError("The blocks do not contain all the parameter indices.")
        if any(int(o) != o or o < 1 for o in self.oversampling):
            raise FlatModelError("Oversampling factors must be integer!")
        for p in np.flatnonzero(self.periodic):
            if not (np.isfinite(self.lower[p]) and np.isfinite(self.upper[p])):
                raise FlatModelError(
                    f"Parameter '{self.names[p]}' cannot be periodic if it is not bounded."
                )
        self.T = None
        if self.proposal_cov is not None:
            self.set_covariance(self.proposal_cov)

    # ---- sizes -----------------------------------------------------------------
    @property
    def D(self) -> int:
        return len(self.names)

    @property
    def n_like(self) -> int:
        return len(self.likes)

    @property
    def n_derived(self) -> int:
        return sum(lk.n_derived for lk in self.likes)

    @property
    def row_width(self) -> int:
        """weight, minuslogpost, sampled, derived, minuslogprior, minuslogprior__0,
        chi2, chi2__<like>... (collection.py:154-159)."""
        return 2 + self.D + self.n_derived + 2 + len(self.ext_priors) + 1 + self.n_like

    def columns(self) -> list:
        cols = ["weight", "minuslogpost"] + list(self.names)
        for lk in self.likes:
            if lk.n_derived:
                names = lk.derived_names or [
                    f"{lk.name}_derived_{i}" for i in range(lk.n_derived)
                ]
                cols += list(names)
        cols += ["minuslogprior", "minuslogprior__0"]
        cols += [f"minuslogprior__{ep.name}" for ep in self.ext_priors]
        cols += ["chi2"]
        cols += [f"chi2__{lk.name}" for lk in self.likes]
        return cols

    # ---- prior -----------------------------------------------------------------
    @property
    def uniform_logp(self) -> float:
        """prior.py:526-532"""
        u = self.prior_kind == PRIOR_UNIFORM
        return float(-np.sum(np.log(self.upper[u] - self.lower[u])))

    # ---- blocking (proposal.py:183-201) ----------------------------------------
    @property
    def i_of_j(self) -> np.ndarray:
        return np.array(list(_chain(*self.blocks)), dtype=np.int32)

    @property
    def block_sizes(self) -> np.ndarray:
        return np.array([len(b) for b in self.blocks], dtype=np.int32)

    @property
    def j_start(self) -> np.ndarray:
        n = self.block_sizes
        return np.array([int(n[:i].sum()) for i in range(len(n))], dtype=np.int32)

    @property
    def last_slow(self) -> int:
        if self.drag and self.i_last_slow_block is not None:
            return int(self.i_last_slow_block)
        return len(self.blocks) - 1

    @property
    def n_slow(self) -> int:
        return int(self.block_sizes[: 1 + self.last_slow].sum())

    @property
    def n_fast(self) -> int:
        return self.D - self.n_slow

    @property
    def cycle_length(self) -> int:
        """mcmc.py:402-407"""
        if self.drag:
            return self.n_slow
        return int(sum(len(b) * o for b, o in zip(self.blocks, self.oversampling)))

    # ---- proposal (proposal.py:226-263) ----------------------------------------
    def set_covariance(self, cov: np.ndarray):
        self.T = transforms_from_cov(cov, self.i_of_j)
        self.proposal_cov = np.array(cov, dtype=np.float64, copy=True)

    def get_covariance(self) -> np.ndarray:
        return self.proposal_cov.copy()

    # ---- convenience constructors ---------------------------------------------
    @classmethod
    def gaussian(cls, means, covs, weights=None, bounds=(-1.0, 1.0), names=None,
                 derived=False, **kw):
        means = np.atleast_2d(np.asarray(means, dtype=np.float64))
        D = means.shape[1]
        names = list(names) if names is not None else [f"x{i}" for i in range(D)]
        lk = LikeSpec.gaussian_mixture(np.arange(D), means, covs, weights, derived=derived,
                                       name="gaussian_mixture")
        b = np.asarray(bounds, dtype=np.float64)
        lower = np.broadcast_to(b[..., 0], (D,)).copy()
        upper = np.broadcast_to(b[..., 1], (D,)).copy()
        return cls(
            names=names, prior_kind=np.zeros(D, np.int32), lower=lower, upper=upper,
            loc=np.zeros(D), pscale=np.ones(D), periodic=np.zeros(D, np.int32),
            likes=[lk], **kw,
        )


def synthetic_gaussian_cov(D: int, seed: int = 20260925):
    """The benchmark target of SURVEY.md section 8d: a correlated Gaussian with
    sigma_i = 0.02 * 10^U(-0.5, 0.5) and a random (A A^T) correlation matrix."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((D, 2 * D))
    C = A @ A.T / (2 * D)
    d = np.sqrt(np.diag(C))
    C = C / d[:, None] / d[None, :]
    sig = 0.02 * 10 ** rng.uniform(-0.5, 0.5, D)
    cov = sig[:, None] * C * sig[None, :]
    return (cov + cov.T) / 2
