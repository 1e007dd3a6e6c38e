"""
Cobaya plugin: ``sampler: mcmc`` executed by the B200 ensemble engine.

This module is imported only when the reference package ``cobaya`` is importable; it is
the drop-in boundary of SURVEY.md section 8b.  ``MCMC`` subclasses
``cobaya.sampler.CovmatSampler`` (sampler.py:467), is constructed by ``cobaya.run.run``
exactly like the reference's ``cobaya.samplers.mcmc.MCMC`` (run.py:161) and honours every
key of ``mcmc.yaml`` (mirrored as class attributes so ``update_info`` accepts them,
input.py:403-434) plus the engine's own keys.

Two ways to select it:

* non-invasive alias:   ``sampler: {cobaya_b200.plugin.MCMC: {...}}``
* true drop-in:         ``cobaya_b200.plugin.install_as_mcmc()`` before ``run(info)``;
                        afterwards ``sampler: mcmc`` resolves to this class.

Host-side setup (start points, blocking, initial covmat) follows ``MCMC.initialize``
(mcmc.py:111-271) through the model's public API; the per-proposal loop runs only on the
GPU.  Models outside the recognised set raise ``LoggedError`` (no CPU fallback).
"""

from __future__ import annotations

import sys
import time
from collections.abc import Callable, Sequence
from itertools import chain
from typing import Any

import os

import numpy as np
import pandas as pd
from cobaya import mpi
from cobaya.collection import SampleCollection, apply_temperature_cov, remove_temperature_cov
from cobaya.conventions import Extension, OutPar, get_version
from cobaya.log import LoggedError
from cobaya.sampler import CovmatSampler
from cobaya.tools import NumberWithUnits, get_external_function
from cobaya.yaml import yaml_dump_file

from .flatmodel import FlatModelError
from .lowering import lower_model
from .mcmc import EnsembleMCMC


class _CurrentPointView:
    """``sampler.current_point`` as the reference's tests/callbacks touch it
    (tests/test_mcmc_initial_covmat.py:87-97): the state of local chain 0."""

    def __init__(self, sampler):
        self._s = sampler
        self.output_thin = 1

    @property
    def values(self):
        if self._s._ens is None:
            return self._s._x0[0]
        return self._s._ens.engine.get_state()["x"][0]

    @property
    def weight(self):
        return 1 if self._s._ens is None else int(self._s._ens.engine.get_state()["weight"][0])

    @property
    def logpost(self):
        if self._s._ens is None:
            return self._s._logpost0[0]
        return float(self._s._ens.engine.get_state()["logpost"][0])


class _ProposerView:
    """``sampler.proposer`` surface used by tests and ``write_checkpoint``."""

    def __init__(self, sampler):
        self._s = sampler

    def get_covariance(self):
        return self._s._fm.get_covariance()

    def set_covariance(self, cov):
        self._s._fm.set_covariance(cov)
        if self._s._ens is not None:
            self._s._ens.engine.set_covariance(cov)

    def get_scale(self):
        return self._s.proposal_scale

    def d(self):
        return self._s._fm.D

    @property
    def i_of_j(self):
        return self._s._fm.i_of_j

    @property
    def j_start(self):
        return list(self._s._fm.j_start)

    @property
    def oversampling_factors(self):
        return np.array(self._s._fm.oversampling, dtype=int)

    @property
    def transform(self):
        T, fm = self._s._fm.T, self._s._fm
        return [T[js:, js: js + n] for js, n in zip(fm.j_start, fm.block_sizes)]


class MCMC(CovmatSampler):
    r"""
    Adaptive, speed-hierarchy-aware MCMC sampler (adapted from CosmoMC)
    \cite{Lewis:2002ah,Lewis:2013hha}, run as an ensemble of lock-step chains on
    NVIDIA B200 GPUs.
    """

    sampler_type: str = "mcmc"
    supports_periodic_params = True
    file_base_name = "mcmc"
    _at_resume_prefer_new = CovmatSampler._at_resume_prefer_new + [
        "burn_in", "callback_function", "callback_every", "max_tries", "output_every",
        "learn_every", "learn_proposal_Rminus1_max", "learn_proposal_Rminus1_max_early",
        "learn_proposal_Rminus1_min", "max_samples", "Rminus1_stop", "Rminus1_cl_stop",
        "Rminus1_cl_level", "covmat", "covmat_params",
    ]
    _at_resume_prefer_old = CovmatSampler._at_resume_prefer_old + ["proposal_scale", "blocking"]

    # ---- every key of cobaya/samplers/mcmc/mcmc.yaml with its default ------------------
    burn_in: Any = 0
    max_tries: Any = "40d"
    covmat: Any = None
    covmat_params: Any = None
    proposal_scale: float = 2.4
    output_every: Any = "60s"
    learn_every: Any = "40d"
    temperature: float | None = 1
    learn_proposal: bool = True
    learn_proposal_Rminus1_max: float = 2.0
    learn_proposal_Rminus1_max_early: float = 30.0
    learn_proposal_Rminus1_min: float = 0.0
    max_samples: float = np.inf
    Rminus1_stop: float = 0.01
    Rminus1_cl_stop: float = 0.2
    Rminus1_cl_level: float = 0.95
    Rminus1_single_split: int = 4
    measure_speeds: Any = True
    oversample_power: float = 0.4
    oversample_thin: Any = True
    drag: bool = False
    blocking: Sequence | None = None
    callback_function: Callable | str | None = None
    callback_every: Any = None
    seed: Any = None
    check_every: Any = None   # deprecated, accepted and ignored (mcmc.yaml:78)
    oversample: Any = None    # deprecated
    drag_limits: Any = None   # deprecated
    # ---- engine keys ---------------------------------------------------------------------
    chains_per_gpu: int = 8192
    device: int | None = None
    rows_per_chain: int | None = None
    launch_cycles: int | None = None
    device_checkpoint: bool | None = None

    # dependency injection for the CPU tests of the host logic (tests/oracle_engine.py);
    # None = the CUDA engine.  Never set by the product.
    _engine_factory = None

    def set_instance_defaults(self):
        super().set_instance_defaults()
        self.converged = False
        self.mpi_size = None
        self.Rminus1_last = np.inf

    # ------------------------------------------------------------------ initialize
    def initialize(self):
        """mcmc.py:111-271, for ``chains_per_gpu`` chains per process."""
        if not self.model.prior.d():
            raise LoggedError(self.log, "No parameters being varied for sampler")
        from . import distributed as cbd

        if not cbd.world_consistent():
            raise LoggedError(
                self.log, "torch.distributed runs %d processes but cobaya.mpi knows of %d: "
                          "call cobaya_b200.distributed.init() on every rank before "
                          "cobaya.run.run(), so that rank-dependent output (chain files, "
                          "root-only checkpoint) is right.", self._torch_world(), mpi.size())
        if self.temperature is None:
            self.temperature = 1
        self._ens = None
        self.n_steps_raw = 0
        self.last_point_callback = 0
        self.i_learn = 1
        self.progress = pd.DataFrame(
            columns=["N", "timestamp", "acceptance_rate", "Rminus1", "Rminus1_cl"])
        if self.callback_function:
            self.callback_function_callable = get_external_function(self.callback_function)
        # Resuming (mcmc.py:131-139,187-214): the ensemble continues from the engine
        # snapshot written next to the chain file at the end of the previous run
        self._resume_snapshot = None
        resuming = bool(self.output and self.output.is_resuming())
        if resuming:
            if max(self.mpi_size or 0, 1) != mpi.size():  # mcmc.py:131-139
                raise LoggedError(
                    self.log, "Cannot resume a run with a different number of chains: "
                              "was %d and now is %d.", max(self.mpi_size or 0, 1), mpi.size())
            try:
                self._resume_snapshot = EnsembleMCMC.load_snapshot(self.snapshot_filename())
            except OSError as e:
                raise LoggedError(
                    self.log, "Cannot resume: engine snapshot '%s' not found (%s). It is "
                              "written at timed outputs and when run() ends.", self.snapshot_filename(), e)
            self.mpi_info("Resuming from previous sample!")
        n_local = int(self.chains_per_gpu)
        # start points: one independent valid point per chain (mcmc.py:215-222)
        D = self.model.prior.d()
        max_tries_init = 10 * D + 100
        self._x0 = np.empty((n_local, D))
        self._logpost0 = np.empty(n_local)
        self.log.info("Getting %d initial points...", n_local)
        if self._resume_snapshot is None:
            if n_local <= self._START_POINTS_ONE_BY_ONE:
                for c in range(n_local):  # exactly the reference's call (mcmc.py:215-222)
                    x, res = self.model.get_valid_point(max_tries=max_tries_init * 100,
                                                        random_state=self._rng)
                    self._x0[c] = x
                    self._logpost0[c] = res.logpost
            else:
                self._x0[:] = self._start_points_batch(n_local, max_tries_init * 100)
                self._logpost0[:] = np.nan  # evaluated on the device by cb2_set_state
        if self._resume_snapshot is not None:
            self._x0[:] = np.nan  # replaced by the snapshot's current points in run()
        if self.measure_speeds and not self.blocking:
            n = None if self.measure_speeds is True else int(self.measure_speeds)
            if not self._measure_speeds_on_device():
                self.model.measure_and_set_speeds(n=n, discard=0, random_state=self._rng)
        self.current_point = _CurrentPointView(self)
        self.set_proposer_blocking()
        self.set_proposer_initial_covmat(load=True)
        # one chain file per process, prefix.<rank+1>.txt (mcmc.py:142-151)
        self.collection = SampleCollection(
            self.model, self.output, name=str(1 + mpi.rank()), temperature=self.temperature,
            sample_type="mcmc", is_batch=True, resuming=resuming)
        self._row_cursor = None   # per-chain rows already handed to the collection
        self._segments = []       # per-drain row counts of every chain (file layout)
        if not resuming:
            self.write_checkpoint()

    # up to this many chains per process the start points come from the reference's own
    # Model.get_valid_point, one call per chain; above, from the vectorised restatement
    _START_POINTS_ONE_BY_ONE = 64

    def _start_points_batch(self, n, max_tries):
        """``n`` independent start points, each drawn like ``Prior.reference`` +
        ``Model.get_valid_point`` draw one (prior.py:866-961, model.py:707-754) -- from the
        ``ref`` pdf (or the fixed ``ref`` value, or the prior where no ``ref`` is given),
        redrawn until the prior density is non-null -- but vectorised over the chains: one
        ``rvs(size=n)`` per parameter instead of n x D scalar calls (the reference takes
        ~2.5 ms per point at D = 64: 20 s for 8192 chains).  For the models the engine
        accepts (lowering.py) the likelihood is finite wherever the prior is, so "valid
        point" reduces to the prior's support; ``cb2_set_state`` then evaluates the
        posterior of every point on the device and refuses non-finite ones."""
        import numbers

        prior, rng = self.model.prior, self._rng
        D = prior.d()
        lower = np.asarray(prior._lower_limits, dtype=np.float64)
        upper = np.asarray(prior._upper_limits, dtype=np.float64)
        from_prior = [i for i, r in enumerate(prior.ref_pdf)
                      if isinstance(r, numbers.Real) and np.isnan(r)]
        if from_prior:
            self.log.info("Reference values or pdfs for some parameters were not provided. "
                          "Sampling from the prior instead for those parameters.")
        out = np.empty((n, D))
        todo = np.arange(n)
        for _ in range(int(max_tries)):
            m = len(todo)
            X = np.empty((m, D))
            for i, ref in enumerate(prior.ref_pdf):
                if hasattr(ref, "rvs"):
                    X[:, i] = ref.rvs(size=m, random_state=rng)
                elif i not in from_prior:
                    X[:, i] = ref
            if from_prior:
                ps = prior.sample(n=m, ignore_external=True, random_state=rng)
                X[:, from_prior] = ps[:, from_prior]
            ok = np.all(np.isfinite(X), axis=1) & np.all(X >= lower, axis=1) & \
                np.all(X <= upper, axis=1)
            out[todo[ok]] = X[ok]
            todo = todo[~ok]
            if not len(todo):
                return out
        raise LoggedError(
            self.log, "Could not sample from the reference pdf a point with non-null prior "
                      "density after %d tries. Maybe your prior is improper of your reference "
                      "pdf is null-defined in the domain of the prior.", max_tries)

    @staticmethod
    def _torch_world():
        # a process group can only exist if torch.distributed has been imported already;
        # importing torch here just to find out that there is none costs seconds
        tdist = sys.modules.get("torch.distributed")
        if tdist is not None and tdist.is_available() and tdist.is_initialized():
            return tdist.get_world_size()
        return 1

    def snapshot_filename(self, rank=None):
        """Engine snapshot of this process (next to ``prefix.<rank+1>.txt``)."""
        import os

        rank = mpi.rank() if rank is None else rank
        return os.path.join(self.output.folder,
                            f"{self.output.prefix}.b200_state.{rank + 1}.npz")

    def rows_filename(self, rank=None):
        """Binary copy of this process' chain file (raw float64 rows, same order): what a
        resumed run reloads into the engine -- the text file keeps ~8 digits."""
        import os

        rank = mpi.rank() if rank is None else rank
        return os.path.join(self.output.folder,
                            f"{self.output.prefix}.b200_rows.{rank + 1}.bin")

    def set_proposer_blocking(self):
        """mcmc.py:320-410: blocks/oversampling from the model, dragging gates, thinning."""
        if self.blocking:
            self.blocks, self.oversampling_factors = self.model.check_blocking(self.blocking)
        else:
            self.blocks, self.oversampling_factors = (
                self.model.get_param_blocking_for_sampler(
                    oversample_power=self.oversample_power, split_fast_slow=self.drag))
        if self.drag:
            if len(self.blocks) == 1:
                self.drag = False
                self.mpi_warning("Dragging disabled: not possible if there is only one block.")
            if max(self.oversampling_factors) / min(self.oversampling_factors) < 2:
                self.drag = False
                self.mpi_warning("Dragging disabled: speed ratios < 2.")
        self.drag_interp_steps = 0
        i_last_slow = None
        if self.drag:
            i_last_slow = next(i for i, o in enumerate(self.oversampling_factors) if o != 1) - 1
            n_slow = len(list(chain(*self.blocks[: 1 + i_last_slow])))
            n_fast = len(list(chain(*self.blocks[1 + i_last_slow:])))
            self.drag_interp_steps = int(
                np.round(self.oversampling_factors[i_last_slow + 1] * n_fast / n_slow))
            if self.drag_interp_steps < 2:
                self.drag = False
                self.mpi_warning("Dragging disabled: "
                                 "speed ratio and fast-to-slow ratio not large enough.")
        output_thin = 1
        if not self.drag and np.any(np.array(self.oversampling_factors) > 1):
            if self.oversample_thin:
                output_thin = int(np.round(
                    sum(len(b) * o for b, o in zip(self.blocks, self.oversampling_factors))
                    / self.model.prior.d()))
        self.current_point.output_thin = output_thin
        self._updated_info["blocking"] = list(zip(self.oversampling_factors, self.blocks))
        sampled = list(self.model.parameterization.sampled_params())
        blocks_indices = [[sampled.index(p) for p in b] for b in self.blocks]
        if self.drag:
            self.cycle_length = sum(len(b) for b in blocks_indices[: 1 + i_last_slow])
        else:
            self.cycle_length = sum(
                len(b) * o for b, o in zip(blocks_indices, self.oversampling_factors))
        self._blocking_lowered = dict(
            blocks=blocks_indices, oversampling=[int(o) for o in self.oversampling_factors],
            drag=bool(self.drag), i_last_slow_block=i_last_slow if self.drag else None,
            drag_interp_steps=self.drag_interp_steps, output_thin=output_thin)
        self.proposer = _ProposerView(self)

    def set_proposer_initial_covmat(self, load=False):
        """mcmc.py:412-440 + lowering of the model (recognised set only)."""
        self._initial_covmat, where_nan = self._load_covmat(
            prefer_load_old=self.output.is_resuming() if self.output else False)
        self._covmat_incomplete = bool(np.any(where_nan))
        try:
            self._fm = lower_model(
                self.model, proposal_cov=apply_temperature_cov(self._initial_covmat,
                                                               self.temperature),
                proposal_scale=float(self.proposal_scale), temperature=float(self.temperature),
                **self._blocking_lowered)
        except FlatModelError as e:
            raise LoggedError(self.log, "%s", str(e)) from e

    def _measure_speeds_on_device(self, n_points=4096, repeats=3):
        """``Model.measure_and_set_speeds`` (model.py:1543-1592) with the components timed where
        they run: every likelihood alone on the device at the start points
        (``cb2_measure_speeds``, CUDA events).  The measured speeds feed the reference's own
        automatic blocking.  Returns False (host measurement) for the CPU test engine or when
        there are no start points yet (resuming)."""
        if self._engine_factory is not None or not np.all(np.isfinite(self._x0)):
            return False
        from .engine import Engine

        d = self.model.prior.d()
        try:
            fm = lower_model(self.model, proposal_cov=np.eye(d))
        except FlatModelError as e:
            raise LoggedError(self.log, "%s", str(e)) from e
        dev = self.device if self.device is not None else int(os.environ.get("LOCAL_RANK", 0))
        try:
            eng = Engine(fm, n_chains=1, seed=0, chain_id0=0, rows_cap=1, device=int(dev))
        except Exception as e:
            raise LoggedError(self.log, "Could not start the B200 engine: %s", e) from e
        try:
            speeds = eng.measure_speeds(self._x0[:n_points], repeats=repeats)
        except Exception as e:
            raise LoggedError(self.log, "Measuring speeds on the device failed: %s", e) from e
        finally:
            eng.close()
        if mpi.more_than_one_process():
            speeds = np.average(mpi.allgather(speeds), axis=0)
        named = dict(zip(self.model.likelihood, speeds))
        self.mpi_info("Setting measured speeds (per sec, on the device): %r",
                      {k: float(f"{v:.3g}") for k, v in named.items()})
        for like, speed in zip(self.model.likelihood.values(), speeds):
            like.set_measured_speed(float(speed))
        return True

    def _check_external_functions(self, ens, n_points=4):
        """An external likelihood runs as CUDA on the device and as Python in the reference's
        ``Model``: evaluate both at a few current points and refuse a disagreement."""
        from .flatmodel import LIKE_EXTERNAL

        if not (any(lk.kind == LIKE_EXTERNAL for lk in self._fm.likes) or self._fm.ext_priors):
            return
        st = ens.engine.get_state()
        for c in range(min(n_points, ens.n_chains_local)):
            want = self.model.logposterior(st["x"][c])  # both untempered
            got = st["logpost"][c]
            if not np.isclose(got, want.logpost, rtol=1e-8, atol=1e-8):
                raise LoggedError(
                    self.log, "The CUDA source of an external likelihood / prior disagrees with its "
                    "Python callable: log-posterior %r (device) vs %r (Python) at %r.",
                    float(got), float(want.logpost), st["x"][c].tolist())

    # ------------------------------------------------------------------ run
    def _options(self):
        keys = ["burn_in", "max_tries", "proposal_scale", "learn_every", "temperature",
                "learn_proposal", "learn_proposal_Rminus1_max",
                "learn_proposal_Rminus1_max_early", "learn_proposal_Rminus1_min",
                "max_samples", "Rminus1_stop", "Rminus1_cl_stop", "Rminus1_cl_level",
                "Rminus1_single_split", "callback_every", "chains_per_gpu", "device",
                "rows_per_chain", "launch_cycles", "device_checkpoint"]
        o = {k: getattr(self, k) for k in keys}
        seed = self.seed
        # one Philox key for the whole run (chains are told apart by their global id); an
        # unseeded run takes it from the root process' generator
        o["seed"] = mpi.share_mpi(int(self._rng.integers(2**62))) if seed is None else int(
            np.random.SeedSequence(seed).generate_state(1, np.uint64)[0] >> 2)
        return o

    def run(self):
        """mcmc.py:451-528 -- the loop body runs on the GPU."""
        with mpi.ProcessState(self):  # a failing rank ends the others too (mcmc.py:469)
            self._run()

    def _run(self):
        dist = None
        if self._torch_world() > 1:
            from .mcmc import TorchDist

            dist = TorchDist()
        try:
            self._ens = EnsembleMCMC(self._fm, self._x0, self._options(), dist=dist,
                                     engine=self._engine_factory,
                                     covmat_incomplete=self._covmat_incomplete,
                                     resume_from=self._restore_rows(self._resume_snapshot))
        except Exception as e:
            raise LoggedError(self.log, "Could not start the B200 engine: %s", e) from e
        ens = self._ens
        self._check_external_functions(ens)
        if self._resume_snapshot is None:
            self._row_cursor = np.zeros(ens.n_chains_local, np.int64)
        self._resume_snapshot = None  # rows of the previous run: free the host copy
        self.mpi_info("Sampling! (%d lock-step chains)", ens.n_chains)

        def _cb(_):
            self.n_steps_raw = ens.n_steps_raw
            if self.callback_function and ens.n() >= self.last_point_callback + max(
                    1, ens.callback_every.value):
                self.callback_function_callable(self)
                self.last_point_callback = ens.n()

        # Timed output (mcmc.py:473-481,696-699): at a convergence check, if `output_every`
        # seconds have passed, the rows stored since the last output leave the device in
        # bulk and are appended to the chain file, and the engine snapshot and
        # .progress/.checkpoint/.covmat are rewritten, so a killed run keeps its chain and
        # resumes from there.  `output_every` without a unit counts accepted steps per chain.
        out_every = NumberWithUnits(self.output_every, "s", dtype=int)
        last_out = {"t": time.time(), "n": 0}

        def _ck(_):
            if not self.output:
                return
            if out_every.unit:
                due = time.time() >= last_out["t"] + out_every.value
            else:
                due = ens.last_summary["min_rows"] >= last_out["n"] + max(1, out_every.value)
            # every process must take the same decision (the snapshots of a run belong together)
            if not mpi.share_mpi(bool(due)):
                return
            self._timed_output(ens)
            last_out["t"] = time.time()
            last_out["n"] = ens.last_summary["min_rows"]

        try:
            ens.run(callback=_cb, on_checkpoint=_ck)
        except LoggedError:
            raise
        except Exception as e:
            # keep what was sampled (a rows-full or device error used to lose it)
            try:
                self._drain_rows()
            except Exception:  # the device may be gone
                pass
            raise LoggedError(self.log, "%s", e) from e
        self.n_steps_raw = ens.n_steps_raw
        self._timed_output(ens)
        self.mpi_info("Sampling complete after %d accepted steps.",
                      ens.last_summary["sum_rows"])

    def _timed_output(self, ens):
        self.converged = ens.converged
        self.Rminus1_last = ens.Rminus1_last
        self._sync_progress(ens)
        self._drain_rows()
        self.write_checkpoint()
        if self.output:
            self._save_state(ens)

    def _sync_progress(self, ens):
        for i, c in enumerate(ens.progress, start=1):
            self.progress.loc[i] = [c.N, c.timestamp, c.acceptance_rate, c.Rminus1,
                                    c.Rminus1_cl]

    def n(self, burn_in=False):
        return 0 if self._ens is None else self._ens.n()

    def _collection_from_rows(self, data, target):
        cols = list(self.collection.columns)
        if data.shape[1] != len(cols):
            raise LoggedError(self.log, "Engine row width %d does not match the collection's "
                                        "%d columns", data.shape[1], len(cols))
        target._data = pd.DataFrame(data, columns=cols, copy=False)  # rows are ours: no copy
        target._cache_reset()
        return target

    # ---- rows: device -> collection -> files ---------------------------------------------
    def _drain_rows(self):
        """Hand the rows stored since the previous drain to ``self.collection`` (a real
        SampleCollection, filled in bulk: SURVEY.md section 8b) and append them to the
        chain file ``prefix.<rank+1>.txt`` through the reference's own writer
        (collection.py:1287-1315) and to its binary twin.

        File layout: every drain appends one *segment* = the new rows of chain 0, of chain
        1, ... of this process.  The file is therefore chronological to within one output
        interval, so that the ``skip``/``thin`` that ``load_samples``/GetDist apply per
        FILE (output.py:324-424) act on every chain's early part, as they do for the
        reference's one-chain-per-file layout."""
        ens = self._ens
        if ens is None or self._row_cursor is None:
            return
        rows, counts = ens.engine.rows_bulk(first=self._row_cursor)
        if not len(rows):
            return
        self._row_cursor = self._row_cursor + counts
        self._segments.append(counts)
        df = pd.DataFrame(rows, columns=list(self.collection.columns), copy=False)
        self.collection._cache_reset()
        self.collection._data = df if not len(self.collection._data) else pd.concat(
            [self.collection._data, df], ignore_index=True)
        if self.output:
            with open(self.rows_filename(), "ab") as f:
                rows.tofile(f)
            self.collection.out_update()

    def _save_state(self, ens):
        """Engine snapshot without rows (they are in the binary rows file, appended to at
        every drain): a few MB however long the run."""
        import os

        snap = ens.snapshot(with_rows=False)
        snap["segments"] = (np.array(self._segments, np.int64).reshape(-1, ens.n_chains_local))
        snap["file_rows"] = np.int64(len(self.collection))
        tmp = self.snapshot_filename() + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, **snap)
        os.replace(tmp, self.snapshot_filename())

    def _restore_rows(self, snap):
        """Resuming: rebuild every chain's rows (chain-major, for the engine) from the
        binary rows file and the segment table of the snapshot, and the collection from
        the same rows (bit-exact; the text file is only checked for length)."""
        if snap is None:
            return None
        import os

        W = int(snap["row_width"])
        seg = np.asarray(snap["segments"], np.int64).reshape(-1, int(snap["n_chains_local"]))
        file_rows = int(snap["file_rows"])
        if seg.sum() != file_rows or not np.array_equal(seg.sum(axis=0), snap["n_rows"]):
            raise LoggedError(self.log, "Cannot resume: the snapshot's segment table does not "
                                        "match its row counts.")
        try:
            flat = np.fromfile(self.rows_filename(), dtype=np.float64, count=file_rows * W)
        except OSError as e:
            raise LoggedError(self.log, "Cannot resume: rows file '%s' not readable (%s)",
                              self.rows_filename(), e)
        if flat.size != file_rows * W:
            raise LoggedError(self.log, "Cannot resume: rows file '%s' holds %d rows, the "
                                        "snapshot expects %d.", self.rows_filename(),
                              flat.size // W, file_rows)
        if os.path.getsize(self.rows_filename()) != file_rows * W * 8:
            with open(self.rows_filename(), "r+b") as f:  # an output the snapshot never saw
                f.truncate(file_rows * W * 8)
        rows = flat.reshape(file_rows, W)
        # segment order -> chain-major order
        n_rows = seg.sum(axis=0)
        chain_off = np.concatenate([[0], np.cumsum(n_rows)[:-1]])
        before = np.zeros_like(n_rows)
        dest = np.empty(file_rows, np.int64)
        pos = 0
        for counts in seg:
            tot = int(counts.sum())
            ids = np.repeat(np.arange(len(counts)), counts)
            starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
            within = np.arange(tot) - np.repeat(starts, counts)
            dest[pos: pos + tot] = chain_off[ids] + before[ids] + within
            before = before + counts
            pos += tot
        chain_major = np.empty_like(rows)
        chain_major[dest] = rows
        # the collection continues the file: same rows, exact values
        n_txt = len(self.collection)
        self._collection_from_rows(rows, self.collection)
        self.collection._n_last_out = file_rows
        if n_txt != file_rows:  # text file ahead of / behind the snapshot: rewrite it
            self.collection._out_delete()
            self.collection._n_last_out = 0
            self.collection.out_update()
        self._segments = [c for c in seg]
        self._row_cursor = n_rows.copy()
        snap = dict(snap)
        snap["rows"] = chain_major
        return snap

    # ------------------------------------------------------------------ products
    def samples(self, combined: bool = False, skip_samples: float = 0,
                to_getdist: bool = False):
        """mcmc.py:1092-1144.  Without arguments: this process' chains, concatenated (the
        collection bound to the output file).  ``combined``: the chains of all processes
        (call it from every process).  ``to_getdist``: one :class:`getdist.MCSamples` built
        by the reference's own ``SampleCollection.to_getdist`` (collection.py:1163-1248)
        from one collection per chain, so GetDist sees the chains separately, as it does the
        reference's MPI chains."""
        if self.temperature != 1 and not to_getdist:
            self.mpi_warning(
                "The MCMC chain(s) are stored with temperature != 1. Keep that in mind when "
                "operating on them, or detemper (in-place) with "
                "products()['sample'].reset_temperature()'.")
        if not (combined or to_getdist):
            if not skip_samples:
                return self.collection
            # skipping is applied to every chain before concatenation (mcmc.py:1127-1143)
            # and returns a copy; the collection bound to the output files is untouched
            return self._collection_from_rows(self._ens.samples(skip_samples=skip_samples),
                                              self.collection.copy(empty=True))
        if not skip_samples:
            self.mpi_warning("When combining chains, it is recommended to remove some "
                             "initial fraction, e.g. 'skip_samples=0.3'")
        ens = self._ens
        if to_getdist:
            try:
                import getdist  # noqa: F401
            except ImportError as e:
                raise LoggedError(self.log, "to_getdist=True needs GetDist (%s)", e) from e
            per_chain = [ens.chain_rows(c, skip_samples) for c in range(ens.n_chains_local)]
            gathered = ens.dist.all_gather_object(per_chain)
            colls = [self._collection_from_rows(rows, self.collection.copy(empty=True))
                     for rank_rows in gathered for rows in rank_rows if len(rows)]
            if not colls:
                raise LoggedError(self.log, "No samples to export.")
            return colls[0].to_getdist(combine_with=colls[1:])
        gathered = ens.dist.all_gather_object(ens.samples(skip_samples=skip_samples))
        return self._collection_from_rows(np.concatenate(gathered),
                                          self.collection.copy(empty=True))

    def products(self, combined: bool = False, skip_samples: float = 0,
                 to_getdist: bool = False) -> dict:
        return {"sample": self.samples(combined, skip_samples, to_getdist),
                "progress": self.progress}

    def write_checkpoint(self):
        """mcmc.py:1045-1078"""
        if mpi.is_main_process() and self.output:
            self.dump_covmat(remove_temperature_cov(self.proposer.get_covariance(),
                                                    self.temperature))
            info = {"sampler": {self.get_name(): {
                "converged": bool(self.converged), "Rminus1_last": self.Rminus1_last,
                "burn_in": 0, "mpi_size": mpi.get_mpi_size()}}}
            yaml_dump_file(self.checkpoint_filename(), info, error_if_exists=False)
            self._write_progress_file()

    def _write_progress_file(self):
        """``prefix.progress`` in the reference's format (mcmc.py:163-181,1067-1077).  The
        reference appends one line per checkpoint; the table is rewritten as a whole here
        (same content), since timed outputs may skip checkpoints."""
        if self.progress is None or self.progress.empty:
            return
        header_fmt = {"N": 6 * " " + "N", "timestamp": 17 * " " + "timestamp"}
        head = "# " + " ".join(header_fmt.get(col, ((7 + 8) - len(col)) * " " + col)
                               for col in self.progress.columns)
        body = self.progress.to_string(header=False, index=False,
                                       formatters={"N": "{:9f}".format})
        tmp = self.progress_filename() + ".tmp"
        with open(tmp, "w", encoding="utf-8") as f:
            f.write(head + "\n" + body + "\n")
        import os
        os.replace(tmp, self.progress_filename())

    def converge_info_changed(self, old_info, new_info):
        keys = ["Rminus1_stop", "Rminus1_cl_stop", "Rminus1_cl_level", "max_samples"]
        return any(old_info.get(p) != new_info.get(p) for p in keys)

    @classmethod
    def output_files_regexps(cls, output, info=None, minimal=False):
        import re

        regexps = [output.collection_regexp(name=None)]
        if minimal:
            return [(r, None) for r in regexps]
        regexps += [re.compile(output.prefix_regexp_str + re.escape(ext.lstrip(".")) + "$")
                    for ext in [Extension.checkpoint, Extension.progress, Extension.covmat]]
        regexps += [re.compile(output.prefix_regexp_str + r"b200_state\.\d+\.npz$")]
        return [(r, None) for r in regexps]

    @classmethod
    def get_version(cls):
        return get_version()

    @classmethod
    def _get_desc(cls, info=None):
        return ("Adaptive, speed-hierarchy-aware MCMC sampler (adapted from CosmoMC) "
                r"\cite{Lewis:2002ah,Lewis:2013hha}, B200 ensemble engine.")


mcmc = MCMC  # ``sampler: {cobaya_b200.plugin.mcmc: ...}`` also resolves


def install_as_mcmc():
    """Make ``sampler: mcmc`` resolve to the B200 engine: the internal lookup imports
    ``cobaya.samplers.mcmc`` through ``importlib`` (tools.py:201), which honours
    ``sys.modules``."""
    import types

    import cobaya.samplers.mcmc as ref  # the reference module (kept for its helpers)

    mod = types.ModuleType("cobaya.samplers.mcmc")
    mod.__dict__.update({k: v for k, v in ref.__dict__.items() if not k.startswith("__")})
    mod.MCMC = MCMC
    mod.mcmc = MCMC
    mod.__package__ = "cobaya.samplers.mcmc"
    mod.__path__ = getattr(ref, "__path__", [])
    mod.__reference_module__ = ref
    sys.modules["cobaya.samplers.mcmc"] = mod
    return mod
