"""
Host-side driver of the ensemble sampler: the part of ``cobaya.samplers.mcmc.MCMC``
that is NOT per-proposal work -- option handling (mcmc.yaml), the outer ``run`` loop
(mcmc.py:451-528), the checkpoint rule (``check_ready`` :752-771), convergence +
covariance learning (``check_convergence_and_learn_proposal`` :773-1032) and products
(:1092-1184) -- restated for M = n_gpus x chains_per_gpu lock-step chains.

Per-proposal work (propose -> logposterior -> accept -> store) happens only in the CUDA
engine (``cobaya_b200.engine.Engine`` -> libcobaya_b200.so).  There is no CPU path.

Ensemble semantics that have no single-chain counterpart in the reference (SURVEY.md
section 7 "hard parts"), chosen to follow the reference's MPI behaviour (one chain per
rank) as closely as lock-step allows:

* checkpoint: taken at the end of the launch in which EVERY chain has at least
  ``i_learn * learn_every`` stored rows (the reference waits until all ranks are ready,
  mcmc.py:500-506, while ready ranks keep sampling); each chain contributes its own
  second half ``[n_c//2:]`` and ``N_c = n_c`` (mcmc.py:787-793);
* ``max_samples``: the run stops when every chain holds at least that many rows;
* a stuck chain (mcmc.py:717-743) is detected at the end of the launch it occurs in.
"""

from __future__ import annotations

import datetime
import os
import logging
import re
from dataclasses import dataclass, field

import numpy as np

from .convergence import rminus1_cl_from_sums, rminus1_from_sums
from .engine import (FLAG_INTERNAL, FLAG_ROWS_FULL, FLAG_STUCK, MOMENTS_HALVES,
                     MOMENTS_SINGLE_SPLIT, Engine, EngineError)
from .flatmodel import FlatModel

log = logging.getLogger("mcmc")


class SamplerError(RuntimeError):
    """Mirror of cobaya.log.LoggedError for the standalone driver."""


class OtherRankError(SamplerError):
    """Another process of the run failed (mirror of cobaya.mpi.OtherProcessError)."""


# defaults of cobaya/samplers/mcmc/mcmc.yaml (every key, verbatim) + the engine's keys
MCMC_DEFAULTS = {
    "burn_in": 0, "max_tries": "40d", "covmat": None, "covmat_params": None,
    "proposal_scale": 2.4, "output_every": "60s", "learn_every": "40d",
    "temperature": 1, "learn_proposal": True, "learn_proposal_Rminus1_max": 2.0,
    "learn_proposal_Rminus1_max_early": 30.0, "learn_proposal_Rminus1_min": 0.0,
    "max_samples": np.inf, "Rminus1_stop": 0.01, "Rminus1_cl_stop": 0.2,
    "Rminus1_cl_level": 0.95, "Rminus1_single_split": 4, "measure_speeds": True,
    "oversample_power": 0.4, "oversample_thin": True, "drag": False, "blocking": None,
    "callback_function": None, "callback_every": None, "seed": None,
    "check_every": None, "oversample": None, "drag_limits": None,
}
ENGINE_DEFAULTS = {
    "chains_per_gpu": 8192,       # lock-step chains on each GPU
    "device": None,               # CUDA device index (default: LOCAL_RANK or 0)
    "rows_per_chain": None,       # sample capacity per chain (default: from max_samples)
    "launch_cycles": None,        # proposal cycles per kernel launch group
    "device_checkpoint": None,    # R-1 / covariance learning on the device (None: when D <= 64)
}


class NumberWithUnits:
    """``"40d"`` -> 40 x scale (mirrors cobaya/tools.py:454-518 for the 'd' and 's' units)."""

    def __init__(self, n_with_unit, unit: str, dtype=int, scale=None):
        self.unit = None
        self.dtype = dtype
        if isinstance(n_with_unit, NumberWithUnits):
            n_with_unit = n_with_unit.original
        self.original = n_with_unit
        if isinstance(n_with_unit, str):
            m = re.fullmatch(r"\s*([0-9.eE+\-]*)\s*(" + unit + r")?\s*", n_with_unit)
            if not m:
                raise SamplerError(f"Could not understand {n_with_unit!r} (unit {unit!r}).")
            num = m.group(1) or "1"
            self.unit_value = dtype(float(num))
            self.unit = m.group(2)
        else:
            self.unit_value = n_with_unit
        self.value = self.unit_value
        if scale is not None:
            self.set_scale(scale)

    def set_scale(self, scale):
        if self.unit:
            self.scale = scale
            self.value = self.dtype(self.unit_value * scale)

    def __bool__(self):
        return bool(self.unit_value)


@dataclass
class Checkpoint:
    N: int
    timestamp: str
    acceptance_rate: float
    Rminus1: float | None
    Rminus1_cl: float | None = None
    learned: bool = False


class _NoDist:
    """Single-process stand-in for the torch.distributed calls used below."""

    rank, size = 0, 1

    def all_reduce_sum(self, arr):
        return arr

    def all_gather_i64(self, vec):
        return [list(vec)]

    def all_gather_object(self, obj):
        return [obj]


class TorchDist:
    """One process per GPU (torchrun): NCCL all-reduce of the per-GPU sufficient statistics
    (replaces mpi.array_gather + share, mcmc.py:791-793,914,1005,1021).  With the gloo
    backend the same code runs on CPU tensors (used by the world_size-2 CPU tests)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self._gather_buf = self._gather_in = None
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl"
            else torch.device("cpu"))

    def buffer(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device=self.device)

    def all_reduce_sum(self, arr):
        t = self.torch.as_tensor(np.asarray(arr, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def all_reduce_tensor_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_object(self, obj):
        out = [None] * self.size
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_gather_i64(self, vec):
        """One small all-gather (the per-launch summary): a single collective and a single
        device->host read instead of three all-reduces with their own copies."""
        tt = self.torch
        n = len(vec)
        if self._gather_buf is None or self._gather_buf.numel() != self.size * n:
            self._gather_buf = tt.zeros(self.size * n, dtype=tt.int64, device=self.device)
            self._gather_in = tt.zeros(n, dtype=tt.int64, device=self.device)
        self._gather_in.copy_(tt.as_tensor(np.asarray(vec, dtype=np.int64)),
                              non_blocking=False)
        self.dist.all_gather_into_tensor(self._gather_buf, self._gather_in, group=self.group)
        return self._gather_buf.cpu().numpy().reshape(self.size, n)


def decide_learning(opts, Rminus1, converged):
    """Learning gate of mcmc.py:1009-1030. Returns (learn?, message)."""
    if not opts["learn_proposal"] or converged:
        return False, ""
    if Rminus1 > opts["learn_proposal_Rminus1_max"]:
        return False, ("Convergence less than requested for updates: "
                       "waiting until the next convergence check.")
    if Rminus1 < opts["learn_proposal_Rminus1_min"]:
        return False, ("Convergence better then better than `learn_proposal_Rminus1_min`"
                       f"={opts['learn_proposal_Rminus1_min']}: covmat will not be updated.")
    return True, " - Updated covariance matrix of proposal pdf."


class EnsembleMCMC:
    """M = world_size x chains_per_gpu lock-step chains of the adaptive Metropolis sampler."""

    def __init__(self, fm: FlatModel, x0: np.ndarray, options: dict | None = None,
                 dist=None, engine: Engine | None = None, covmat_incomplete: bool = False,
                 resume_from: dict | None = None):
        opts = dict(MCMC_DEFAULTS)
        opts.update(ENGINE_DEFAULTS)
        unknown = set(options or {}) - set(opts)
        if unknown:
            raise SamplerError(f"Unknown mcmc option(s): {sorted(unknown)}")  # input.py:403
        opts.update(options or {})
        self.opts = opts
        self.fm = fm
        self.dist = dist or _NoDist()
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, fm.D)
        self.n_chains_local = x0.shape[0]
        self.n_chains = self.n_chains_local * self.dist.size
        if opts["temperature"] is None:
            opts["temperature"] = 1
        if opts["temperature"] < 1:
            log.warning("Sampling temperatures <1 can lead to innacurate inference.")
        fm.temperature = float(opts["temperature"])
        fm.proposal_scale = float(opts["proposal_scale"])
        self.temperature = fm.temperature
        # quantities in units of the cycle length (mcmc.py:122-125,409-410)
        self.cycle_length = fm.cycle_length
        self.output_thin = int(fm.output_thin)
        scale = self.cycle_length // self.output_thin
        if opts["callback_every"] is None:
            opts["callback_every"] = opts["learn_every"]
        self.max_tries = NumberWithUnits(opts["max_tries"], "d", scale=scale,
                                         dtype=float)
        self.learn_every = NumberWithUnits(opts["learn_every"], "d", scale=scale)
        self.callback_every = NumberWithUnits(opts["callback_every"], "d", scale=scale)
        self.burn_in = NumberWithUnits(opts["burn_in"], "d", scale=scale)
        mt = self.max_tries.value
        fm.max_tries = int(min(mt, 2**59)) if np.isfinite(mt) else 2**59
        self.max_samples = opts["max_samples"]
        if covmat_incomplete and opts["learn_proposal"]:  # mcmc.py:419-429
            opts["learn_proposal_Rminus1_max"] = opts["learn_proposal_Rminus1_max_early"]
        lc = opts["launch_cycles"]
        if lc is None:
            lc = max(1, (self.learn_every.value // 2) // max(self.cycle_length, 1))
        self.launch_steps = int(lc) * self.cycle_length
        # Sample capacity per chain: the INITIAL size of the device row store.  The store
        # grows on demand (_ensure_row_capacity), so neither a finite `max_samples` (the
        # run stops when the SLOWEST chain has that many rows; faster chains hold more) nor
        # `max_samples: inf` needs a worst-case allocation up front.  Without
        # `rows_per_chain` the engine fits the wanted size into 40 % of the free memory.
        rows = opts["rows_per_chain"]
        rows_want = (int(self.max_samples) + self.launch_steps + 2 * self.learn_every.value
                     if np.isfinite(self.max_samples) else 16 * self.learn_every.value)
        self.rows_per_chain = None if rows is None else int(rows)
        seed = opts["seed"]
        self.seed = int(seed) if seed is not None else int(
            np.random.SeedSequence().generate_state(2, np.uint32).view(np.uint64)[0] >> 1)
        device = opts["device"]
        if device is None:
            import os

            device = int(os.environ.get("LOCAL_RANK", "0"))
        snap = resume_from
        if snap is not None:
            self._check_snapshot(snap)
            self.seed = int(snap["seed"])
            fm.set_covariance(np.asarray(snap["proposal_cov"]))  # the learned proposal
            if "proposal_T" in snap:
                # the transform exactly as the stopped run used it: the device checkpoint
                # (cb2_checkpoint_device) and LAPACK agree to rounding only, and a resumed run
                # continues bit for bit
                fm.T = np.array(snap["proposal_T"], dtype=np.float64, copy=True)
            need = int(np.max(snap["n_rows"])) + self.launch_steps
            rows_want = max(rows_want, need + 2 * self.learn_every.value)
            if self.rows_per_chain is not None:
                self.rows_per_chain = max(self.rows_per_chain, need)
        # `engine`: an engine object, or a factory with Engine's signature (dependency
        # injection for the CPU tests of this host logic; the default is the CUDA engine and
        # nothing else is ever chosen here)
        if engine is None or callable(engine):
            engine = (engine or Engine)(
                fm, n_chains=self.n_chains_local, seed=self.seed, device=int(device),
                chain_id0=self.dist.rank * self.n_chains_local, rows_cap=self.rows_per_chain,
                burn_in=int(self.burn_in.value), rows_want=rows_want)
        self.engine = engine
        self.rows_per_chain = self.engine.rows_cap
        if snap is not None and self.rows_per_chain < int(np.max(snap["n_rows"])):
            raise SamplerError(
                "Cannot resume: the device cannot hold the %d rows per chain of the "
                "snapshot (room for %d)." % (int(np.max(snap["n_rows"])), self.rows_per_chain))
        self.converged = False
        self.Rminus1_last = np.inf
        self.i_learn = 1
        self.n_steps_raw = 0
        self.progress: list[Checkpoint] = []
        if snap is None:
            self.engine.set_state(x0)
        else:
            self.engine.import_state(
                snap["blob"], np.asarray(snap["rows"]).reshape(-1, fm.row_width),
                np.asarray(snap["n_rows"], np.int64))
            self.Rminus1_last = float(snap["Rminus1_last"])
            self.i_learn = int(snap["i_learn"])
            self.n_steps_raw = int(snap["n_steps_raw"])
            self.progress = [
                Checkpoint(N=int(r[0]), timestamp=str(t), acceptance_rate=float(r[1]),
                           Rminus1=None if np.isnan(r[2]) else float(r[2]),
                           Rminus1_cl=None if np.isnan(r[3]) else float(r[3]),
                           learned=bool(r[4]))
                for r, t in zip(np.asarray(snap["progress"]).reshape(-1, 5),
                                snap["progress_timestamps"])]
        self._shift = np.asarray(self.dist.all_reduce_sum(x0.sum(axis=0))) / self.n_chains
        if snap is not None:
            self._shift = np.asarray(snap["shift"], dtype=np.float64)
        self._mom_buf = None
        self.last_summary = None
        self.local_summary = None

    # ------------------------------------------------------------------ resuming
    SNAPSHOT_VERSION = 1

    def snapshot(self, with_rows: bool = True) -> dict:
        """Everything needed to continue this rank's chains in a new process: the engine's
        per-chain state and stored rows plus the driver's checkpoint bookkeeping
        (the ensemble counterpart of mcmc.py:187-214,1045-1078)."""
        eng = self.engine
        if with_rows:
            rows_all, n_rows = eng.rows_bulk()
        else:
            rows_all = np.zeros((0, self.fm.row_width))
            n_rows = eng.get_state()["n_rows"]
        prog = np.array([[c.N, c.acceptance_rate,
                          np.nan if c.Rminus1 is None else c.Rminus1,
                          np.nan if c.Rminus1_cl is None else c.Rminus1_cl,
                          float(c.learned)] for c in self.progress], dtype=np.float64)
        return dict(
            version=self.SNAPSHOT_VERSION, blob=eng.export_state(),
            n_rows=np.asarray(n_rows, np.int64), rows=rows_all,
            proposal_cov=self.fm.get_covariance(), proposal_T=np.asarray(self.fm.T),
            seed=np.uint64(self.seed),
            rank=self.dist.rank, world=self.dist.size, n_chains_local=self.n_chains_local,
            D=self.fm.D, row_width=self.fm.row_width,
            converged=bool(self.converged), Rminus1_last=float(self.Rminus1_last),
            i_learn=int(self.i_learn), n_steps_raw=int(self.n_steps_raw),
            shift=np.asarray(self._shift, dtype=np.float64),
            progress=prog.reshape(-1, 5),
            progress_timestamps=np.array([c.timestamp for c in self.progress], dtype=str),
        )

    def save_snapshot(self, path: str):
        """Written to a temporary name and renamed: a crash never leaves a torn snapshot."""
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, **self.snapshot())
        os.replace(tmp, path)

    @staticmethod
    def load_snapshot(path: str) -> dict:
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}

    def _check_snapshot(self, snap):
        if int(snap["version"]) != self.SNAPSHOT_VERSION:
            raise SamplerError("Snapshot written by another version of the engine.")
        if int(snap["world"]) != self.dist.size or int(snap["rank"]) != self.dist.rank:
            raise SamplerError(  # mcmc.py:131-139
                "Cannot resume a run with a different number of chains: was %d processes "
                "and now is %d." % (int(snap["world"]), self.dist.size))
        if int(snap["n_chains_local"]) != self.n_chains_local or int(snap["D"]) != self.fm.D \
                or int(snap["row_width"]) != self.fm.row_width:
            raise SamplerError(
                "Cannot resume: the snapshot holds %d chains of a %d-parameter model "
                "(%d columns), the input asks for %d chains, %d parameters (%d columns)." % (
                    int(snap["n_chains_local"]), int(snap["D"]), int(snap["row_width"]),
                    self.n_chains_local, self.fm.D, self.fm.row_width))

    # ------------------------------------------------------------------ helpers
    def _global_summary(self, local_error: str | None = None):
        """The 8-word engine summary of every rank in ONE collective (an all-gather of 9
        int64 per rank; min, max and sums are then taken on the host).  The ninth word
        carries "this rank failed", so that a rank-local error (CUDA error, device memory)
        is raised on every rank in the same iteration instead of leaving the others blocked
        in the next collective (the role of ProcessState, mpi.py:350-467)."""
        if local_error is None:
            s = self.engine.summary()
        else:
            s = dict(min_rows=0, max_rows=0, sum_rows=0, n_stuck=0, n_rows_full=0,
                     n_internal=0, sum_accepted=0, sum_weight=0)
        self.local_summary = s
        vec = [s["min_rows"], s["max_rows"], s["sum_rows"], s["n_stuck"], s["n_rows_full"],
               s["n_internal"], s["sum_accepted"], s["sum_weight"], int(local_error is not None)]
        allv = np.asarray(self.dist.all_gather_i64(vec), dtype=np.int64).reshape(-1, 9)
        if allv[:, 8].any():
            bad = [int(r) for r in np.nonzero(allv[:, 8])[0]]
            if local_error is not None:
                raise SamplerError(local_error)
            raise OtherRankError(f"Another process failed (rank {bad}) - exiting.")
        g = dict(min_rows=int(allv[:, 0].min()), max_rows=int(allv[:, 1].max()),
                 sum_rows=int(allv[:, 2].sum()), n_stuck=int(allv[:, 3].sum()),
                 n_rows_full=int(allv[:, 4].sum()), n_internal=int(allv[:, 5].sum()),
                 sum_accepted=int(allv[:, 6].sum()), sum_weight=int(allv[:, 7].sum()))
        self.last_summary = g
        return g

    def _ensure_row_capacity(self):
        """Grow the device row store before a chain can run out of room: one launch adds at
        most ``launch_steps`` rows per chain (the reference's collection grows without
        limit, collection.py:765-778)."""
        eng = self.engine
        have = (self.local_summary or eng.summary())["max_rows"]
        if have + self.launch_steps <= eng.rows_cap:
            return
        need = have + self.launch_steps
        want = max(need, int(1.5 * eng.rows_cap) + self.launch_steps)
        free, _, held = eng.mem_info()
        per_row = self.n_chains_local * self.fm.row_width * 8
        fit = int(0.9 * free) // per_row  # old and new store coexist during the re-layout
        new_cap = min(want, fit)
        if new_cap < need:
            raise SamplerError(
                "Device memory exhausted by the stored samples: %d chains x %d rows hold "
                "%.1f GB and %.1f GB are free; lower chains_per_gpu or use output_thin / "
                "max_samples." % (self.n_chains_local, eng.rows_cap, held / 1e9, free / 1e9))
        eng.grow_rows(new_cap)
        self.rows_per_chain = eng.rows_cap
        log.info("Sample store grown to %d rows per chain (%.1f GB).", eng.rows_cap,
                 eng.rows_cap * per_row / 1e9)

    def n(self):
        """Stored rows of the shortest chain (the ensemble analogue of MCMC.n())."""
        return (self.last_summary or self._global_summary())["min_rows"]

    def _check_health(self, g):
        if g["n_internal"]:
            raise SamplerError("internal engine error (basis/tape window violated)")
        if g["n_stuck"]:  # mcmc.py:720-743
            raise SamplerError(
                "The chain has been stuck for %d attempts, stopping sampling "
                "(%d of %d chains). Make sure the reference point is sensible and initial "
                "covmat." % (int(self.max_tries.value), g["n_stuck"], self.n_chains))
        if g["n_rows_full"]:
            raise SamplerError(
                f"{g['n_rows_full']} chains filled their sample storage "
                f"(rows_per_chain={self.rows_per_chain}); increase rows_per_chain.")

    # ------------------------------------------------------------------ run loop
    def run(self, callback=None, on_checkpoint=None):
        """MCMC.run (mcmc.py:451-528) for the ensemble.  ``on_checkpoint(self)`` is called
        after every convergence check (where the reference writes its checkpoint,
        mcmc.py:1029-1032) -- the plugin uses it for the timed output of mcmc.py:473-481."""
        log.info("Sampling! (%d chains on %d GPU(s))", self.n_chains, self.dist.size)
        g = self._global_summary()
        while g["min_rows"] < self.max_samples and not self.converged:
            err = None
            try:
                self._ensure_row_capacity()
                self.engine.advance(self.launch_steps)
            except Exception as e:  # rank-local: every rank must leave the loop together
                err = f"{type(e).__name__}: {e}"
            self.n_steps_raw += self.launch_steps
            g = self._global_summary(local_error=err)
            self._check_health(g)
            if callback is not None:
                callback(self)
            if self.check_ready(g):
                self.check_convergence_and_learn_proposal()
                self.i_learn += 1
                if on_checkpoint is not None:
                    on_checkpoint(self)
        if g["min_rows"] >= self.max_samples:
            log.info("Reached maximum number of accepted steps allowed (%s). Stopping.",
                     self.max_samples)
        log.info("Sampling complete after %d accepted steps.", g["sum_rows"])
        return self

    def check_ready(self, g=None):
        """mcmc.py:752-771 with the ensemble rule: all chains reached the next multiple."""
        g = g or self._global_summary()
        return g["min_rows"] >= self.i_learn * self.learn_every.value and g["min_rows"] > 0

    def _moments(self, mode, split):
        eng, d = self.engine, self.dist
        if isinstance(d, TorchDist) and d.backend == "nccl":
            if self._mom_buf is None:
                self._mom_buf = d.buffer(eng.moments_len)
            eng.moments(mode=mode, split=split, shift=self._shift,
                        dev_ptr=self._mom_buf.data_ptr(), host=False)
            eng.sync()
            d.all_reduce_tensor_(self._mom_buf)      # NCCL over NVLink
            return self._mom_buf.cpu().numpy()
        local = eng.moments(mode=mode, split=split, shift=self._shift)
        return np.asarray(d.all_reduce_sum(local))

    def _use_device_checkpoint(self):
        opt = self.opts.get("device_checkpoint")
        able = (self.n_chains > 1 and self.fm.D <= 64 and
                hasattr(self.engine, "checkpoint_device") and
                (isinstance(self.dist, _NoDist) or getattr(self.dist, "backend", "") == "nccl"))
        if opt and not able:
            raise SamplerError("device_checkpoint needs D <= 64, more than one chain and the "
                               "CUDA engine (single process or NCCL).")
        return able if opt is None else bool(opt)

    def _checkpoint_device(self):
        """Sums -> (NCCL all-reduce) -> cb2_checkpoint_device; nothing D x D leaves the GPU."""
        eng, d = self.engine, self.dist
        if isinstance(d, TorchDist):
            if self._mom_buf is None:
                self._mom_buf = d.buffer(eng.moments_len)
            eng.moments(mode=MOMENTS_HALVES, split=0, shift=self._shift,
                        dev_ptr=self._mom_buf.data_ptr(), host=False)
            eng.sync()
            d.all_reduce_tensor_(self._mom_buf)
            d.torch.cuda.current_stream().synchronize()
            return eng.checkpoint_device(dev_ptr=self._mom_buf.data_ptr())
        eng.moments(mode=MOMENTS_HALVES, split=0, shift=self._shift, host=False)
        return eng.checkpoint_device()

    def _bounds(self, mode, split, limfrac):
        """Per-chain confidence bounds -> all-reduced sums (mcmc.py:918-939)."""
        eng, d = self.engine, self.dist
        n = 1 + 4 * self.fm.D
        if isinstance(d, TorchDist) and d.backend == "nccl":
            if self._mom_buf is None:
                self._mom_buf = d.buffer(eng.moments_len)
            buf = self._mom_buf[:n]
            eng.bounds(limfrac, mode=mode, split=split, shift=self._shift,
                       dev_ptr=buf.data_ptr(), host=False)
            eng.sync()
            d.all_reduce_tensor_(buf)
            return buf.cpu().numpy()
        local = eng.bounds(limfrac, mode=mode, split=split, shift=self._shift)
        return np.asarray(d.all_reduce_sum(local))

    def check_convergence_and_learn_proposal(self):
        """mcmc.py:773-1032 on all-reduced sums; identical result on every rank."""
        o = self.opts
        on_device = self._use_device_checkpoint()
        if on_device:
            # D x D algebra and the new transform on the GPU (cb2_checkpoint_device)
            res = self._checkpoint_device()
        else:
            if self.n_chains > 1:
                sums = self._moments(MOMENTS_HALVES, 0)
            else:
                try:
                    sums = self._moments(MOMENTS_SINGLE_SPLIT, int(o["Rminus1_single_split"]))
                except Exception as e:  # mcmc.py:816-821
                    log.info("Not enough points in chain to check convergence. (%s)", e)
                    return
            res = rminus1_from_sums(sums, self.fm.D, self._shift)
        cp = Checkpoint(N=res["N"], timestamp=datetime.datetime.now().isoformat(),
                        acceptance_rate=float(res["acceptance"]), Rminus1=res["Rminus1"])
        self.progress.append(cp)
        log.info(" - Acceptance rate: %.3f", res["acceptance"])
        if not res["success"]:
            log.warning("Negative covariance eigenvectors or failed eigenvalues. "
                        "Skipping learning a new covmat for now.")  # mcmc.py:872-887
            return
        Rminus1 = res["Rminus1"]
        log.info(" - Convergence of means: R-1 = %f after %d accepted steps", Rminus1,
                 res["N"])
        converged_means = max(Rminus1, self.Rminus1_last) < o["Rminus1_stop"]  # :908
        if converged_means:
            # R-1 of the confidence-interval bounds (mcmc.py:918-1002): per-chain weighted
            # quantiles at Rminus1_cl_level/2 on the device, std over chains / sqrt(diag W)
            try:
                mode = MOMENTS_HALVES if self.n_chains > 1 else MOMENTS_SINGLE_SPLIT
                bsums = self._bounds(mode, int(o["Rminus1_single_split"]),
                                     float(o["Rminus1_cl_level"]) / 2.0)
                if "W" not in res:
                    res["W"] = self.engine.checkpoint_cov()
                Rminus1_cl = rminus1_cl_from_sums(bsums, self.fm.D, res["W"])
            except Exception as e:  # mcmc.py:936-938,999-1002
                log.info("Computation of the bounds was not possible (%s). "
                         "Waiting until the next converge check.", e)
                Rminus1_cl = None
            if Rminus1_cl is not None:
                cp.Rminus1_cl = Rminus1_cl
                log.info(" - Convergence of bounds: R-1 = %f after %d accepted steps",
                         Rminus1_cl, res["N"])
                if Rminus1_cl < o["Rminus1_cl_stop"]:  # mcmc.py:994-996
                    self.converged = True
                    log.info("The run has converged!")
        self.Rminus1_last = Rminus1
        self._shift = np.asarray(res["mean"], dtype=np.float64)
        learn, msg = decide_learning(o, Rminus1, self.converged)
        if msg:
            log.info(msg)
        if learn:
            try:
                if on_device and res["proposal_ok"]:
                    try:
                        self.engine.adopt_proposal(res.get("W"))  # repacked on the device
                    except EngineError:
                        # constant blocks that only the host packer builds
                        self.engine.set_covariance(self.engine.checkpoint_cov())
                else:
                    if "W" not in res:
                        res["W"] = self.engine.checkpoint_cov()
                    self.engine.set_covariance(res["W"])  # is already tempered (mcmc.py:1023)
                cp.learned = True
            except Exception as e:
                log.debug("Updating covariance matrix failed unexpectedly (%s); "
                          "waiting until next covmat learning attempt.", e)

    # ------------------------------------------------------------------ products
    def chain_rows(self, chain: int, skip: float = 0):
        """Rows (SampleCollection layout, collection.py:154-159) of one local chain."""
        rows = self.engine.rows(chain)
        k = int(skip * len(rows)) if 0 < skip < 1 else int(skip)
        return rows[k:]

    def samples(self, chains=None, skip_samples: float = 0.0, return_counts: bool = False):
        """Concatenated rows of the selected local chains (default: all), each with its
        first ``skip_samples`` fraction/number of rows removed (mcmc.py:1127-1143).  All
        chains / a contiguous range leave the device in bulk (cb2_copy_rows_bulk)."""
        eng = self.engine
        if chains is None:
            chains = range(self.n_chains_local)
        chains = list(chains)
        contiguous = chains == list(range(chains[0], chains[0] + len(chains))) if chains else False
        if not contiguous:
            per = [self.chain_rows(c, skip_samples) for c in chains]
            out = np.concatenate(per) if per else np.zeros((0, self.fm.row_width))
            return (out, np.array([len(r) for r in per], np.int64)) if return_counts else out
        n_rows = eng.get_state()["n_rows"][chains[0]: chains[-1] + 1]
        if 0 < skip_samples < 1:
            first = (skip_samples * n_rows).astype(np.int64)
        else:
            first = np.minimum(np.full_like(n_rows, int(skip_samples)), n_rows)
        rows, counts = eng.rows_bulk(first=first, chains=(chains[0], chains[-1] + 1))
        return (rows, counts) if return_counts else rows

    def products(self, skip_samples: float = 0.0, chains=None):
        import pandas as pd

        data = self.samples(chains, skip_samples)
        return {
            "sample": pd.DataFrame(data, columns=self.fm.columns(), copy=False),
            "progress": pd.DataFrame(
                [dict(N=c.N, timestamp=c.timestamp, acceptance_rate=c.acceptance_rate,
                      Rminus1=c.Rminus1, Rminus1_cl=c.Rminus1_cl) for c in self.progress]),
        }

    def mean_and_cov(self):
        """Ensemble posterior mean and covariance from the second halves of all chains
        (the checkpoint statistics): cov = W + B."""
        sums = self._moments(MOMENTS_HALVES if self.n_chains > 1 else MOMENTS_SINGLE_SPLIT,
                             int(self.opts["Rminus1_single_split"]))
        res = rminus1_from_sums(sums, self.fm.D, self._shift)
        cov = res["W"] + (res.get("B", 0.0) if self.n_chains > 1 else 0.0)
        return res["mean"], cov, res
