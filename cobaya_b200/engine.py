"""
Thin object wrapper over the C ABI (include/cobaya_b200.h): one ``Engine`` = one GPU =
``n_chains`` lock-step chains.  No numerics happen here; numpy arrays are passed to the
library as plain pointers.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .flatmodel import (LIKE_CONSTANT, LIKE_EXTERNAL, LIKE_GAUSSIAN_MIXTURE, LIKE_ROSENBROCK,
                        FlatModel)

FLAG_STUCK = 1
FLAG_ROWS_FULL = 2
FLAG_INTERNAL = 4
MOMENTS_HALVES = 0
MOMENTS_SINGLE_SPLIT = 1
# `max_tries: inf` is sent as this cap: the stuck test multiplies it by 10 during burn-in
# (mcmc.py:717-719) and must stay inside int64
MAX_TRIES_CAP = 2**59


class EngineError(RuntimeError):
    pass


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    def __init__(self, fm: FlatModel, n_chains: int, seed: int, device: int = 0,
                 chain_id0: int = 0, rows_cap: int | None = 1024, burn_in: int = 0,
                 rows_want: int = 4096, mem_fraction: float = 0.4):
        """``rows_cap``: stored rows per chain the engine can hold before ``grow_rows``;
        ``None`` sizes it to ``rows_want``, limited to ``mem_fraction`` of the free device
        memory (the store can grow later, up to what the device holds)."""
        self.lib = _cabi.load()
        self.fm = fm
        self.n_chains = int(n_chains)
        self.D = fm.D
        self.seed = int(seed)
        self.chain_id0 = int(chain_id0)
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.cb2_create(self.device, self.n_chains, self.D, self.seed,
                                 self.chain_id0, C.byref(h))
        if rc != 0:
            raise EngineError(self.lib.cb2_last_error(None).decode())
        self.h = h
        self._pinned = []
        if rows_cap is None:
            free, _, _ = self.mem_info()
            fit = int(mem_fraction * free / (self.n_chains * fm.row_width * 8))
            rows_cap = max(16, min(int(rows_want), fit))
        self.rows_cap = int(rows_cap)
        self.burn_in = int(burn_in)
        self._upload_model()

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.cb2_last_error(self.h).decode() or f"error {rc}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb2_destroy(self.h)
            self.h = None
        for p in getattr(self, "_pinned", []):
            self.lib.cb2_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _upload_model(self):
        fm, p = self.fm, _cabi.ptr
        self._ck(self.lib.cb2_set_prior(
            self.h, p(_i32(fm.prior_kind)), p(_f64(fm.lower)), p(_f64(fm.upper)),
            p(_f64(fm.loc)), p(_f64(fm.pscale)), p(_i32(fm.periodic)), fm.uniform_logp))
        if np.any(fm.prior_kind >= 2):
            self._ck(self.lib.cb2_set_prior_shapes(
                self.h, p(_f64(fm.pa)), p(_f64(fm.pb)), p(_f64(fm.prior_log_norm))))
        self._ck(self.lib.cb2_clear_likelihoods(self.h))
        for lk in fm.likes:
            if lk.kind == LIKE_GAUSSIAN_MIXTURE:
                self._ck(self.lib.cb2_add_gaussian_mixture(
                    self.h, lk.dim, p(_i32(lk.idx)), lk.n_modes, p(_f64(lk.means)),
                    p(_f64(lk.linv)), p(_f64(lk.logdet)), p(_f64(lk.weights)),
                    int(lk.derived)))
            elif lk.kind == LIKE_ROSENBROCK:
                self._ck(self.lib.cb2_add_rosenbrock(self.h, lk.dim, p(_i32(lk.idx)),
                                                     float(lk.scale)))
            elif lk.kind == LIKE_EXTERNAL:
                self._ck(self.lib.cb2_add_external_likelihood(
                    self.h, lk.dim, p(_i32(lk.idx)), lk.source.encode(), lk.fn_name.encode()))
            elif lk.kind == LIKE_CONSTANT:
                self._ck(self.lib.cb2_add_constant(self.h, float(lk.scale)))
            else:
                raise EngineError(f"unknown likelihood kind {lk.kind}")
        for ep in getattr(fm, "ext_priors", []):
            self._ck(self.lib.cb2_add_external_prior(
                self.h, ep.dim, p(_i32(ep.idx)), ep.source.encode(), ep.fn_name.encode()))
        self._ck(self.lib.cb2_set_blocking(
            self.h, len(fm.blocks), p(_i32(fm.block_sizes)), p(_i32(fm.oversampling)),
            p(_i32(fm.i_of_j)), int(fm.drag), int(fm.last_slow),
            int(fm.drag_interp_steps)))
        if fm.T is not None:
            self.set_proposal(fm.T, fm.proposal_scale)
        self._ck(self.lib.cb2_set_options(
            self.h, float(fm.temperature), self.burn_in,
            int(min(fm.max_tries, MAX_TRIES_CAP)), int(fm.output_thin), self.rows_cap))

    # ------------------------------------------------------------------ API
    def set_proposal(self, T, proposal_scale=None):
        """BlockedProposer.set_covariance result (proposal.py:226-260)."""
        scale = self.fm.proposal_scale if proposal_scale is None else proposal_scale
        self._ck(self.lib.cb2_set_proposal(self.h, _cabi.ptr(_f64(T)), float(scale)))

    def set_covariance(self, cov):
        self.fm.set_covariance(cov)
        self.set_proposal(self.fm.T)

    def set_state(self, x0):
        x0 = _f64(x0).reshape(self.n_chains, self.D)
        self._ck(self.lib.cb2_set_state(self.h, _cabi.ptr(x0)))

    def get_state(self):
        n, D = self.n_chains, self.D
        x = np.empty((n, D)); lp = np.empty(n)
        w = np.empty(n, np.int64); nr = np.empty(n, np.int64); na = np.empty(n, np.int64)
        fl = np.empty(n, np.uint32)
        p = _cabi.ptr
        self._ck(self.lib.cb2_get_state(self.h, p(x), p(lp), p(w), p(nr), p(na), p(fl)))
        return dict(x=x, logpost=lp, weight=w, n_rows=nr, n_accepted=na, flags=fl)

    def logpost(self, X):
        X = _f64(np.atleast_2d(X))
        n = X.shape[0]
        NL, ND = self.fm.n_like, self.fm.n_derived
        lp = np.empty(n); pr = np.empty(n); ll = np.empty((n, NL))
        der = np.empty((n, max(ND, 1)))
        p = _cabi.ptr
        self._ck(self.lib.cb2_logpost(self.h, p(X), n, p(lp), p(pr), p(ll), p(der)))
        return lp, pr, ll, der[:, :ND]

    def advance(self, n_proposals: int):
        self._ck(self.lib.cb2_advance(self.h, int(n_proposals)))

    def sync(self):
        self._ck(self.lib.cb2_sync(self.h))

    def summary(self):
        out = np.zeros(8, np.int64)
        self._ck(self.lib.cb2_summary(self.h, _cabi.ptr(out)))
        keys = ["min_rows", "max_rows", "sum_rows", "n_stuck", "n_rows_full", "n_internal",
                "sum_accepted", "sum_weight"]
        return dict(zip(keys, (int(v) for v in out)))

    @property
    def moments_len(self):
        return 3 + self.D + 2 * self.D * self.D

    def moments(self, mode=MOMENTS_HALVES, split=4, shift=None, dev_ptr=None, host=True):
        out = np.empty(self.moments_len) if host else None
        sh = None if shift is None else _f64(shift)
        self._ck(self.lib.cb2_moments(self.h, int(mode), int(split), _cabi.ptr(sh),
                                      C.c_void_p(dev_ptr) if dev_ptr else None,
                                      _cabi.ptr(out)))
        return out

    def measure_speeds(self, X, repeats=3):
        """Evaluations per second of every likelihood component on the device
        (Model.measure_and_set_speeds, model.py:1543-1592)."""
        X = _f64(np.atleast_2d(X))
        out = np.zeros(self.fm.n_like)
        self._ck(self.lib.cb2_measure_speeds(self.h, _cabi.ptr(X), X.shape[0], int(repeats),
                                             _cabi.ptr(out)))
        return out

    # ---- device checkpoint (D <= 64): R-1, W and the new transform without host LAPACK
    def checkpoint_device(self, dev_ptr=None):
        """cb2_checkpoint_device on the sums of the last ``moments`` call (``dev_ptr``: the
        all-reduced buffer).  Returns the dict of ``convergence.rminus1_from_sums`` without
        W (``checkpoint_cov`` fetches it when somebody needs it on the host)."""
        out = np.empty(8 + self.D)
        self._ck(self.lib.cb2_checkpoint_device(
            self.h, C.c_void_p(dev_ptr) if dev_ptr else None, _cabi.ptr(out)))
        ok = bool(out[5] > 0.5) and bool(np.isfinite(out[3]))
        return dict(M=int(round(out[0])), N=int(round(out[1])), acceptance=float(out[2]),
                    Rminus1=float(out[3]) if ok else None, success=ok,
                    proposal_ok=bool(out[4] > 0.5), sweeps=int(out[6]), mean=out[8:].copy())

    def checkpoint_cov(self):
        W = np.empty((self.D, self.D))
        self._ck(self.lib.cb2_checkpoint_cov(self.h, _cabi.ptr(W)))
        return W

    def adopt_proposal(self, cov=None):
        """The candidate transform of the last ``checkpoint_device`` becomes the proposal
        (device-side repack); the host model follows with the transform read back."""
        self._ck(self.lib.cb2_adopt_proposal(self.h))
        self.fm.T = self.get_proposal()
        self.fm.proposal_cov = self.checkpoint_cov() if cov is None else cov

    def get_proposal(self):
        T = np.empty((self.D, self.D))
        self._ck(self.lib.cb2_get_proposal(self.h, _cabi.ptr(T)))
        return T

    def bounds(self, limfrac, mode=MOMENTS_HALVES, split=4, shift=None, dev_ptr=None,
               host=True):
        """Sums of the per-chain confidence bounds (mcmc.py:918-1002): [1 + 4D]."""
        out = np.empty(1 + 4 * self.D) if host else None
        sh = None if shift is None else _f64(shift)
        self._ck(self.lib.cb2_bounds(self.h, int(mode), int(split), float(limfrac),
                                     _cabi.ptr(sh), C.c_void_p(dev_ptr) if dev_ptr else None,
                                     _cabi.ptr(out)))
        return out

    def rows(self, chain: int, first: int = 0, n: int | None = None):
        """Rows [first, first + n) of one chain (all stored rows by default)."""
        W = self.lib.cb2_row_width(self.h)
        if n is None:
            one = np.zeros(1, np.int64)
            f = np.array([int(first)], np.int64)
            n = self.lib.cb2_copy_rows_bulk(self.h, int(chain), int(chain) + 1, _cabi.ptr(f),
                                            _cabi.ptr(one), None, 0)
            if n < 0:
                raise EngineError(self.lib.cb2_last_error(self.h).decode())
        out = np.empty((max(int(n), 1), W))
        got = self.lib.cb2_copy_rows(self.h, int(chain), int(first), int(n), _cabi.ptr(out))
        if got < 0:
            raise EngineError(self.lib.cb2_last_error(self.h).decode())
        return out[:got]

    # budget of one bulk transfer (rows are staged once on the device before they leave it)
    BULK_BYTES = 2 << 30

    def rows_bulk(self, first=None, chains=None):
        """Rows ``[first[c], n_rows[c])`` of the chains ``chains = (begin, end)`` (default:
        all), chain-major in one array, and the per-chain counts.  One device-side
        compaction and one D2H copy per ~2 GB (cb2_copy_rows_bulk), instead of one copy per
        chain -- the bulk counterpart of SampleCollection's per-row append
        (collection.py:402-427)."""
        W = self.lib.cb2_row_width(self.h)
        c0, c1 = (0, self.n_chains) if chains is None else (int(chains[0]), int(chains[1]))
        n = c1 - c0
        counts = np.zeros(n, np.int64)
        f = None if first is None else np.ascontiguousarray(first, dtype=np.int64)
        if f is not None and f.shape != (n,):
            raise EngineError("rows_bulk: `first` needs one entry per selected chain")
        total = self.lib.cb2_copy_rows_bulk(self.h, c0, c1, _cabi.ptr(f), _cabi.ptr(counts),
                                            None, 0)
        if total < 0:
            raise EngineError(self.lib.cb2_last_error(self.h).decode())
        out = np.empty((int(total), W))
        budget = max(1, self.BULK_BYTES // (8 * W))
        ends = np.cumsum(counts)
        a = 0
        while a < n:  # chain ranges of at most `budget` rows (at least one chain)
            base = ends[a - 1] if a else 0
            b = int(np.searchsorted(ends, base + budget, side="right"))
            b = min(max(b, a + 1), n)
            r0, r1 = int(base), int(ends[b - 1])
            if r1 > r0:
                fa = None if f is None else np.ascontiguousarray(f[a:b])
                got = self.lib.cb2_copy_rows_bulk(self.h, c0 + a, c0 + b, _cabi.ptr(fa), None,
                                                  _cabi.ptr(out[r0:r1]), r1 - r0)
                if got != r1 - r0:
                    raise EngineError(self.lib.cb2_last_error(self.h).decode()
                                      or "rows changed during the bulk copy")
            a = b
        return out, counts

    def mem_info(self):
        """(free, total, held by the row store) bytes of the engine's GPU."""
        v = np.zeros(3, np.int64)
        self._ck(self.lib.cb2_mem_info(self.h, _cabi.ptr(v[0:1]), _cabi.ptr(v[1:2]),
                                       _cabi.ptr(v[2:3])))
        return int(v[0]), int(v[1]), int(v[2])

    def grow_rows(self, new_cap: int):
        """Larger per-chain sample capacity; stored rows are kept (cb2_grow_rows)."""
        self._ck(self.lib.cb2_grow_rows(self.h, int(new_cap)))
        self.rows_cap = max(self.rows_cap, int(new_cap))

    # ---- asynchronous drain (run loop) ---------------------------------------------------
    def host_buffer(self, n_doubles: int):
        """Page-locked float64 buffer (cb2_host_alloc) as a numpy array; freed with the
        engine (or explicitly by ``free_host_buffer``)."""
        n_doubles = max(int(n_doubles), 1)
        p = self.lib.cb2_host_alloc(n_doubles * 8)
        if not p:
            raise EngineError(f"cannot page-lock {n_doubles * 8 / 1e9:.2f} GB of host memory")
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n_doubles,))
        self._pinned.append(p)
        return arr

    def drain_start(self, out: np.ndarray, counts: np.ndarray | None = None) -> int:
        """Start handing the rows added since the previous drain to ``out`` (flat or
        [rows, W], ideally from ``host_buffer``); returns the number of rows on their way.
        The copy runs on its own stream under the next ``advance``; read ``out`` only after
        ``drain_wait``."""
        W = self.lib.cb2_row_width(self.h)
        got = self.lib.cb2_drain_start(self.h, _cabi.ptr(out), out.size // W, _cabi.ptr(counts))
        if got < 0:
            raise EngineError(self.lib.cb2_last_error(self.h).decode())
        return int(got)

    def drain_wait(self):
        self._ck(self.lib.cb2_drain_wait(self.h))

    def drain_reset(self, first=None):
        f = None if first is None else np.ascontiguousarray(first, dtype=np.int64)
        self._ck(self.lib.cb2_drain_reset(self.h, _cabi.ptr(f)))

    # ---- resuming ---------------------------------------------------------------------
    def export_state(self) -> np.ndarray:
        """Opaque per-chain sampler state (cb2_export_state) as a uint8 array."""
        n = int(self.lib.cb2_snapshot_size(self.h))
        if n < 0:
            raise EngineError("cb2_snapshot_size failed")
        buf = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_export_state(self.h, _cabi.ptr(buf), n))
        return buf

    def import_state(self, blob, rows=None, counts=None):
        """Restore a snapshot taken from an identically configured engine; ``rows`` is the
        list of per-chain row arrays (``self.rows(c)`` of the exporting engine), or, with
        ``counts``, all rows chain-major in one array (``rows_bulk()`` of the exporter)."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._ck(self.lib.cb2_import_state(self.h, _cabi.ptr(blob), blob.size))
        if rows is not None and counts is not None:
            counts = np.ascontiguousarray(counts, dtype=np.int64)
            if counts.shape != (self.n_chains,):
                raise EngineError("import_state: one count per chain is required")
            W = self.lib.cb2_row_width(self.h)
            rows = _f64(rows).reshape(-1, W)
            if rows.shape[0] != counts.sum():
                raise EngineError("import_state: rows do not match the counts")
            if counts.size and counts.max() > self.rows_cap:
                raise EngineError("import_state: more rows per chain than rows_cap")
            ends = np.cumsum(counts)
            budget = max(1, self.BULK_BYTES // (8 * W))
            a = 0
            while a < self.n_chains:
                base = ends[a - 1] if a else 0
                b = int(np.searchsorted(ends, base + budget, side="right"))
                b = min(max(b, a + 1), self.n_chains)
                r0, r1 = int(base), int(ends[b - 1])
                if r1 > r0:
                    self._ck(self.lib.cb2_load_rows_bulk(
                        self.h, a, b, _cabi.ptr(np.ascontiguousarray(counts[a:b])),
                        _cabi.ptr(rows[r0:r1])))
                a = b
        elif rows is not None:
            if len(rows) != self.n_chains:
                raise EngineError("import_state: one row array per chain is required")
            keep = []
            for c, r in enumerate(rows):
                r = _f64(r)
                keep.append(r)
                self._ck(self.lib.cb2_load_rows(self.h, c, r.shape[0] if r.size else 0,
                                                _cabi.ptr(r) if r.size else None))
            self.sync()

    def debug_basis(self, chain, block, epoch):
        n = int(self.fm.block_sizes[block])
        R = np.empty((n, n))
        self._ck(self.lib.cb2_debug_basis(self.h, int(chain), int(block), int(epoch),
                                          _cabi.ptr(R)))
        return R

    def launch_count(self):
        return int(self.lib.cb2_launch_count(self.h))

    def timer_start(self):
        self._ck(self.lib.cb2_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.cb2_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def last_step_kernel(self):
        return int(self.lib.cb2_last_step_kernel(self.h))

    def window_counts(self, reset: bool = False):
        """Windows run by each step kernel, and windows that left their preferred kernel."""
        v = np.zeros(8, np.int64)
        self._ck(self.lib.cb2_window_counts(self.h, _cabi.ptr(v), int(reset)))
        names = ["general", "dmma", "dmma-producer-consumer", "dmma-streamed",
                 "pc_launch_refused", "streamed_did_not_fit", "of_pc_split_products",
                 "chunked_over_chains"]
        return {k: int(x) for k, x in zip(names, v)}

    def ext_route_counts(self):
        """Route of the windows of a model with external functions (cb2_ext_route_counts)."""
        v = np.zeros(4, np.int64)
        self._ck(self.lib.cb2_ext_route_counts(self.h, _cabi.ptr(v)))
        return dict(fused_windows=int(v[0]), graph_replays=int(v[1]),
                    fused_loaded=bool(v[2]), fused_failed=bool(v[3]))

    def debug_message(self):
        return self.lib.cb2_debug_message(self.h).decode()

    def set_kernel_policy(self, policy: int):
        self._ck(self.lib.cb2_set_kernel_policy(self.h, int(policy)))

    def set_profiling(self, on: bool):
        self._ck(self.lib.cb2_set_profiling(self.h, int(bool(on))))

    def kernel_times(self, reset: bool = True):
        """Per-kernel-class device time (ms) and launch counts: tape, basis, step, moments,
        normals (the latter run on a side stream under the step kernel)."""
        ms = np.zeros(5); n = np.zeros(5, np.int64)
        self._ck(self.lib.cb2_kernel_times(self.h, _cabi.ptr(ms), _cabi.ptr(n), int(reset)))
        names = ["tape", "basis", "step", "moments", "normals"]
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}
