"""
ctypes front-end of the C parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product never does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmcmc_oracle.so")
MAX_BLOCKS = 16


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mcmc_oracle.c")
    hdr = os.path.join(_HERE, "mcmc_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH)
        for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Like(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("n_modes", C.c_int32),
        ("derived", C.c_int32),
        ("idx", C.c_void_p), ("means", C.c_void_p), ("linv", C.c_void_p),
        ("logdet", C.c_void_p), ("weights", C.c_void_p), ("scale", C.c_double),
        ("fn", C.c_void_p),
    ]


class _Model(C.Structure):
    _fields_ = [
        ("D", C.c_int32),
        ("prior_kind", C.c_void_p), ("lower", C.c_void_p), ("upper", C.c_void_p),
        ("loc", C.c_void_p), ("pscale", C.c_void_p), ("pa", C.c_void_p), ("pb", C.c_void_p),
        ("periodic", C.c_void_p),
        ("uniform_logp", C.c_double),
        ("n_like", C.c_int32), ("likes", C.c_void_p),
        ("n_blocks", C.c_int32),
        ("block_size", C.c_int32 * MAX_BLOCKS), ("oversampling", C.c_int32 * MAX_BLOCKS),
        ("i_of_j", C.c_void_p),
        ("drag", C.c_int32), ("i_last_slow_block", C.c_int32),
        ("drag_interp_steps", C.c_int32),
        ("T", C.c_void_p), ("proposal_scale", C.c_double),
        ("temperature", C.c_double), ("max_tries", C.c_int64), ("output_thin", C.c_int32),
        ("n_ext_prior", C.c_int32), ("ext_priors", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_row_width.restype = C.c_int32
        L.orc_n_derived.restype = C.c_int32
        L.orc_logpost.restype = C.c_double
        L.orc_chain_new.restype = C.c_void_p
        L.orc_chain_new.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                    C.c_int64]
        L.orc_chain_free.argtypes = [C.c_void_p]
        L.orc_chain_set_model.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_chain_advance.restype = C.c_int
        L.orc_chain_advance.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        C.c_void_p]
        L.orc_chain_get.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_accept_exp.restype = C.c_double
        L.orc_accept_exp.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_radial.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64,
                                 C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_random_SO_N.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32,
                                      C.c_uint32, C.c_void_p]
        L.orc_basis_normals.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32,
                                        C.c_uint32, C.c_void_p]
        L.orc_so_n_from_normals.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_permutation.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32,
                                      C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        L.orc_ensemble_advance.restype = C.c_int
        L.orc_ensemble_advance.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_int32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_HOST_PRELUDE = ("#include <cmath>\n#include <cstring>\n#define __device__\n"
                 "#define __forceinline__ inline\n#define __restrict__\n"
                 "static inline double __longlong_as_double(long long x) "
                 "{ double d; std::memcpy(&d, &x, 8); return d; }\n")
_host_functions = {}


def host_function(source: str, name: str, keep=None):
    """Address of the external likelihood function ``name`` of ``source`` -- the CUDA source
    the engine compiles with NVRTC -- compiled as HOST code with g++ (the CUDA qualifiers
    defined away), so that the oracle evaluates the same arithmetic (test infrastructure)."""
    import hashlib
    import subprocess
    import tempfile

    key = hashlib.sha1((source + "\0" + name).encode()).hexdigest()
    if key not in _host_functions:
        d = tempfile.mkdtemp(prefix="cb2_oracle_ext_")
        cpp, so = os.path.join(d, "f.cpp"), os.path.join(d, "f.so")
        with open(cpp, "w") as f:
            f.write(_HOST_PRELUDE + source)
        res = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("host build of the external function failed:\n" + res.stderr)
        _host_functions[key] = C.CDLL(so)
    lib = _host_functions[key]
    if keep is not None:
        keep.append(lib)
    return C.cast(getattr(lib, name), C.c_void_p).value


class OracleModel:
    """Keeps the numpy arrays alive behind an ``orc_model`` struct."""

    def __init__(self, fm):
        self.fm = fm
        self._keep = []
        k = self._keepa
        m = _Model()
        m.D = fm.D
        m.prior_kind = _p(k(fm.prior_kind, np.int32))
        m.lower = _p(k(fm.lower)); m.upper = _p(k(fm.upper))
        m.loc = _p(k(fm.loc)); m.pscale = _p(k(fm.pscale))
        m.pa = _p(k(fm.pa)); m.pb = _p(k(fm.pb))
        m.periodic = _p(k(fm.periodic, np.int32))
        m.uniform_logp = fm.uniform_logp
        likes = (_Like * max(1, fm.n_like))()
        for i, lk in enumerate(fm.likes):
            likes[i].kind = lk.kind
            likes[i].dim = lk.dim
            likes[i].n_modes = lk.n_modes
            likes[i].derived = int(lk.derived)
            likes[i].idx = _p(k(lk.idx, np.int32))
            if lk.means is not None:
                likes[i].means = _p(k(lk.means)); likes[i].linv = _p(k(lk.linv))
                likes[i].logdet = _p(k(lk.logdet)); likes[i].weights = _p(k(lk.weights))
            likes[i].scale = lk.scale
            if lk.kind == 3:   # external function: the engine's CUDA source compiled for the host
                likes[i].fn = host_function(lk.source, lk.fn_name, keep=self._keep)
        self._likes = likes
        m.n_like = fm.n_like
        m.likes = C.cast(likes, C.c_void_p)
        m.n_blocks = len(fm.blocks)
        for b, (bl, o) in enumerate(zip(fm.blocks, fm.oversampling)):
            m.block_size[b] = len(bl)
            m.oversampling[b] = int(o)
        m.i_of_j = _p(k(fm.i_of_j, np.int32))
        m.drag = int(fm.drag)
        m.i_last_slow_block = fm.last_slow
        m.drag_interp_steps = int(fm.drag_interp_steps)
        m.T = _p(k(fm.T))
        m.proposal_scale = fm.proposal_scale
        m.temperature = fm.temperature
        m.max_tries = int(min(fm.max_tries, 2**59))  # x10 during burn-in must fit int64
        m.output_thin = int(fm.output_thin)
        eps = list(getattr(fm, "ext_priors", []) or [])
        if eps and fm.drag:
            raise ValueError("external priors are not supported together with dragging")
        ext = (_Like * max(1, len(eps)))()
        for i, ep in enumerate(eps):
            ext[i].kind = 3
            ext[i].dim = ep.dim
            ext[i].idx = _p(k(ep.idx, np.int32))
            ext[i].fn = host_function(ep.source, ep.fn_name, keep=self._keep)
        self._ext_priors = ext
        m.n_ext_prior = len(eps)
        m.ext_priors = C.cast(ext, C.c_void_p)
        self.c = m

    def _keepa(self, a, dtype=np.float64):
        a = np.ascontiguousarray(a, dtype=dtype)
        self._keep.append(a)
        return a

    @property
    def ptr(self):
        return C.byref(self.c)

    @property
    def width(self):
        return lib().orc_row_width(self.ptr)

    def logpost(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lp = C.c_double()
        ll = np.zeros(max(1, self.fm.n_like))
        der = np.zeros(max(1, self.fm.n_derived))
        v = lib().orc_logpost(self.ptr, _p(x), C.byref(lp), _p(ll), _p(der))
        return v, lp.value, ll[: self.fm.n_like], der[: self.fm.n_derived]


class OracleChain:
    def __init__(self, om: OracleModel, seed: int, chain_id: int, x0, burn_in: int = 0):
        self.om = om
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.h = lib().orc_chain_new(om.ptr, seed, chain_id, _p(x0), burn_in)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_chain_free(self.h)
            self.h = None

    def set_model(self, om: OracleModel):
        self.om = om
        lib().orc_chain_set_model(self.h, om.ptr)

    def advance(self, n, rows_cap=None):
        """Returns (rc, rows[n_rows, width])."""
        cap = int(rows_cap if rows_cap is not None else n)
        W = self.om.width
        rows = np.zeros((max(cap, 1), W))
        nr = C.c_int64(0)
        rc = lib().orc_chain_advance(self.h, int(n), _p(rows), cap, C.byref(nr))
        return rc, rows[: nr.value].copy()

    def state(self):
        D = self.om.fm.D
        x = np.zeros(D)
        lp = C.c_double(); w = C.c_int64(); ns = C.c_int64(); na = C.c_int64()
        bl = C.c_int64()
        lib().orc_chain_get(self.h, _p(x), C.byref(lp), C.byref(w), C.byref(ns),
                            C.byref(na), C.byref(bl))
        return dict(x=x, logpost=lp.value, weight=w.value, n_steps=ns.value,
                    n_accepted=na.value, burn_in_left=bl.value)


def ensemble_advance(chains, n_proposals, rows_cap=0, n_threads=0, store=True):
    """Advance a list of OracleChain (OpenMP over chains)."""
    n = len(chains)
    arr = (C.c_void_p * n)(*[c.h for c in chains])
    W = chains[0].om.width
    n_rows = np.zeros(n, dtype=np.int64)
    rows = np.zeros((n, max(rows_cap, 1), W)) if store else None
    rc = lib().orc_ensemble_advance(arr, n, int(n_proposals),
                                    _p(rows) if store else None, int(rows_cap),
                                    _p(n_rows), int(n_threads))
    return rc, rows, n_rows


# ---- unit-level helpers ------------------------------------------------------
def philox4x32(key, ctr):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(key[0], key[1], ctr[0], ctr[1], ctr[2], ctr[3], _p(out))
    return out


def random_SO_N(n, seed, chain_id, block, epoch):
    R = np.zeros((n, n))
    lib().orc_random_SO_N(n, seed, chain_id, block, epoch, _p(R))
    return R


def basis_normals(n, seed, chain_id, block, epoch):
    nn = (n + 2) * (n - 1) // 2
    xx = np.zeros(nn + 2)
    lib().orc_basis_normals(n, seed, chain_id, block, epoch, _p(xx))
    return xx[:nn].copy()


def so_n_from_normals(n, xx):
    xx = np.ascontiguousarray(np.concatenate([xx, [0.0, 0.0]]), dtype=np.float64)
    R = np.zeros((n, n))
    lib().orc_so_n_from_normals(n, _p(xx), _p(R))
    return R


def permutation(sorted_idx, seed, chain_id, which, cycle):
    s = np.ascontiguousarray(sorted_idx, dtype=np.int32)
    out = np.zeros_like(s)
    lib().orc_permutation(len(s), seed, chain_id, which, cycle, _p(s), _p(out))
    return out


def radial(n_block, seed, chain_id, t, sub=0):
    r = C.c_double(); s = C.c_double()
    lib().orc_radial(n_block, seed, chain_id, t, sub, C.byref(r), C.byref(s))
    return r.value, s.value


def accept_exp(seed, chain_id, t, sub=0):
    return lib().orc_accept_exp(seed, chain_id, t, sub)


# ---- checkpoint statistics, restating mcmc.py:773-889 --------------------------
def chain_window_stats(rows, D, first, last=None):
    """SampleCollection.mean/cov (collection.py:893-981) over rows[first:last]:
    weighted mean, np.cov(ddof=0, fweights) and acceptance (mcmc.py:311-318)."""
    r = rows[first:last]
    w = r[:, 0]
    X = r[:, 2 : 2 + D]
    mean = np.average(X.T, weights=w, axis=-1)
    cov = np.atleast_2d(np.cov(X.T, ddof=0, fweights=w.astype(np.int64)))
    acc = len(r) / w.sum()
    return mean, cov, acc


def rminus1_from_chain_stats(Ns, means, covs):
    """Root-side block of mcmc.py:856-889.  Returns (Rminus1, mean_of_covs)."""
    from scipy.linalg import lapack

    Ns = np.asarray(Ns, dtype=np.float64)
    mean_of_covs = np.average(covs, weights=Ns, axis=0)           # :856
    cov_of_means = np.atleast_2d(np.cov(np.asarray(means).T))     # :860
    d = np.sqrt(np.diag(cov_of_means))                            # :864
    corr_of_means = (cov_of_means / d).T / d                      # :865
    norm_mean_of_covs = (mean_of_covs / d).T / d                  # :866
    chol = np.linalg.cholesky(norm_mean_of_covs)                  # :871
    Linv = lapack.dtrtri(chol, lower=True)[0]
    eigvals = np.linalg.eigvalsh(Linv.dot(corr_of_means).dot(Linv.T))  # :881
    return float(max(np.abs(eigvals))), mean_of_covs              # :889
