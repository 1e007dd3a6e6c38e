#!/usr/bin/env python
"""
bench.py -- proposals/s of the adaptive-Metropolis hot path on BASELINE.json configs[1]:
64-D correlated Gaussian, 8192 lock-step chains per B200, proposal-covariance learning on.

    python bench.py --gpus N --steps K --warmup W            (N=1; N>1 under torchrun)
    python bench.py --impl reference --steps K --warmup W    (reference on host cores)

A "step" is one pass of the hot path: every chain of every GPU makes ``--locksteps``
proposals (default 16 proposal cycles = 1024 lock-steps), followed by the run loop's own
bookkeeping (summary read-back, and the convergence-check + covariance-learning
checkpoint whenever every chain has accumulated ``learn_every`` more rows -- with an NCCL
all-reduce of the per-GPU moments when N>1).  ``value`` is device-timed (CUDA events on
the engine's stream) with all inputs resident in HBM; ``e2e`` times the same pass through
the public host API with pinned-host inputs uploaded and results read back every step.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 64
CHAINS_PER_GPU = 8192
WORKLOAD = "64-D correlated Gaussian, 8192 chains/GPU, covmat learning on (BASELINE configs[1])"
METRIC = "mcmc_proposals_per_sec"
UNIT = "proposals/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


FP64_NOMINAL_TFLOPS = 37.0  # 148 SM x 64 FP64 FMA/clk x 2 x 1.965 GHz (nominal, HGX B200)


def fp64_peak():
    """FP64 datapath peak: tools/fp64_peak.cu (m8n8k4 DMMA, 32 warps/SM, 8 independent
    accumulators) measured on this pool's B200s; nominal figure if the record is missing."""
    p = os.path.join(ROOT, "profiles", "r1_fp64_peak.json")
    try:
        with open(p) as f:
            return float(json.load(f)["mma_m8n8k4_ilp8_w32"]["tflops"]), \
                "measured (tools/fp64_peak.cu, profiles/r1_fp64_peak.json)"
    except Exception:
        return FP64_NOMINAL_TFLOPS, "nominal"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.mark0 = self.mark1 = None

    def mark(self, which):
        """Remember how many samples existed when the timed region started / ended."""
        if which == 0:
            self.mark0 = len(self.rows)
        else:
            self.mark1 = len(self.rows)

    def run(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.p.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        try:
            self.p.terminate()
        except Exception:
            pass
        # samples taken inside the timed region; the region can be shorter than nvidia-smi's
        # start-up, so the sampler runs from before the warm-up (the same kernels, the same
        # load) and falls back to those samples when the region itself holds fewer than 3
        rows, where = self.rows, "warm-up + timed region"
        if self.mark0 is not None and self.mark1 is not None and self.mark1 - self.mark0 >= 3:
            rows, where = self.rows[self.mark0:self.mark1], "timed region"
        elif self.mark0 is not None and len(self.rows) - self.mark0 >= 3:
            rows, where = self.rows[self.mark0:], "timed region + e2e leg"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown",
                                    "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled_during": where}


# ------------------------------------------------------------------ CPU legs
def cpu_oracle_baseline(fm, prob, target_seconds=12.0):
    """The C oracle (a scalar port of the reference algorithm) on the host cores, on a
    bounded sample of the same workload."""
    from oracle import oracle as orc

    cores = os.cpu_count() or 1
    om = orc.OracleModel(fm)
    x0 = prob.start(4 * cores, 999)
    chains = [orc.OracleChain(om, 1, 10**6 + c, x0[c]) for c in range(len(x0))]
    cyc = max(1, fm.cycle_length)
    t = time.perf_counter()
    orc.ensemble_advance(chains, cyc, store=False, n_threads=cores)
    dt = time.perf_counter() - t
    rate = len(chains) * cyc / dt
    n = int(max(cyc, min(20000, target_seconds * rate / len(chains))) // cyc * cyc)
    t = time.perf_counter()
    orc.ensemble_advance(chains, n, store=False, n_threads=cores)
    dt = time.perf_counter() - t
    return {"value": len(chains) * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(chains)} chains x {n} proposals of the same target ({prob.key}; "
                      f"oracle/mcmc_oracle.c, OpenMP over chains), {dt:.1f} s"}


_REF_WORKER = r"""
import os, sys, time, json
import numpy as np
root, n_samples, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sys.path.insert(0, os.path.join(root, "oracle", "shims"))
sys.path.insert(0, os.path.join(root, "baseline", "_ref"))
sys.path.insert(0, root)
os.environ["COBAYA_NOMPI"] = "1"
import logging
logging.disable(logging.CRITICAL)
from cobaya_b200.flatmodel import synthetic_gaussian_cov
from cobaya.model import get_model
from cobaya.sampler import get_sampler
D = 64
cov = synthetic_gaussian_cov(D)
names = [f"x{i}" for i in range(D)]
info = {"likelihood": {"gaussian_mixture": {"means": [np.zeros(D)], "covs": [cov],
                                            "input_params": names, "derived": False}},
        "params": {n: {"prior": {"min": -1, "max": 1},
                       "ref": {"dist": "norm", "loc": 0, "scale": 0.001}} for n in names},
        "sampler": {"mcmc": {"covmat": np.diag(np.diag(cov)), "covmat_params": names,
                             "measure_speeds": False, "learn_proposal": True, "burn_in": 0,
                             "seed": seed, "max_samples": n_samples, "Rminus1_stop": 1e-9,
                             "output_every": "1000s"}}}
out = []
for rep in range(int(sys.argv[4])):
    model = get_model(info)
    sampler = get_sampler(info["sampler"], model)
    t = time.perf_counter()
    sampler.run()
    dt = time.perf_counter() - t
    out.append((int(sampler.n_steps_raw), dt))
print(json.dumps(out))
"""


def reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "cobaya"))


def run_reference_arm(args):
    """bench.py --impl reference: the UNMODIFIED reference (cobaya 3.6.2 installed in
    baseline/_ref) on the host cores: one independent single-chain process per core
    (the reference's MPI mode is one chain per rank; mpi4py is not installed), each step a
    bounded run of ``max_samples`` accepted rows.  Falls back to the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cobaya_b200 import problems

    # the CPU legs evaluate the same function through the built-in Rosenbrock of the oracle
    prob = problems.get("c3" if args.config == "c3x" else args.config)
    fm = prob.fm
    cores = os.cpu_count() or 1
    steps, warm = args.steps, args.warmup
    peak, _ = peaks()
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "gpu_launches": 0,
            "config": {"workload": prob.workload, "D": fm.D}}
    if reference_available() and prob.key == "c1":
        n_samples = 250  # accepted rows per chain per step (about 0.15 s of run())
        procs = []
        t0 = time.perf_counter()
        # one chain per core, one thread per process (what an `mpirun -np <cores>` run of the
        # reference uses): keep BLAS / numba from oversubscribing the cores
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1",
                   MKL_NUM_THREADS="1", NUMBA_NUM_THREADS="1")
        for c in range(cores):
            procs.append(subprocess.Popen(
                [sys.executable, "-c", _REF_WORKER, ROOT, str(n_samples), str(100 + c),
                 str(steps + warm)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True, env=env))
        res = []
        for p in procs:
            o, _ = p.communicate()
            try:
                res.append(json.loads(o.strip().splitlines()[-1]))
            except Exception:
                pass
        wall = time.perf_counter() - t0
        if res:
            res = np.array(res, dtype=float)[:, warm:, :]  # [proc, step, (n, dt)]
            n_tot = res[:, :, 0].sum()
            t_run = res[:, :, 1].sum(axis=1).max()       # slowest process' run() time
            value = n_tot / t_run
            base.update(value=value, ms_per_step=1e3 * t_run / steps,
                        cpu_baseline={"value": value, "unit": UNIT, "cores": len(res),
                                      "kind": "reference",
                                      "sample": f"{len(res)} independent cobaya 3.6.2 "
                                                f"processes (1 chain each, numba on), "
                                                f"{steps} runs of max_samples={n_samples}; "
                                                f"wall {wall:.0f}s"},
                        e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0})
            emit(base)
            return
    # configs[2]/[3] (and a box without the reference package): the oracle port
    cb = cpu_oracle_baseline(fm, prob, target_seconds=max(5.0, 2.0 * steps))
    base.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                     "d2h_bytes_per_step": 0})
    emit(base)


# ------------------------------------------------------------------ GPU arm
KERNEL_NAMES = {0: "general", 1: "dmma", 2: "dmma-producer-consumer", 3: "dmma-streamed"}


def e2e_through_cobaya_run(chains, max_samples=200):
    """The README quick start, timed: ``cobaya.run.run(info)`` with ``sampler: mcmc`` resolved
    to the engine (install_as_mcmc), 64-D target, ``chains_per_gpu`` chains, no output
    prefix; wall clock around run(): model construction, 8192 start points drawn through the
    reference's own ``Model.get_valid_point``, sampling, bulk hand-over of every stored row
    into a real SampleCollection.  Needs the reference package (baseline/_ref)."""
    if not reference_available():
        return {"unavailable": "reference package not installed under baseline/_ref"}
    for p_ in (os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "baseline", "_ref")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    import logging

    logging.disable(logging.CRITICAL)
    try:
        from cobaya.run import run

        import cobaya_b200.plugin as plugin
        from cobaya_b200.flatmodel import synthetic_gaussian_cov

        cov = synthetic_gaussian_cov(D)
        names = [f"x{i}" for i in range(D)]
        info = {"likelihood": {"gaussian_mixture": {"means": [np.zeros(D)], "covs": [cov],
                                                    "input_params": names, "derived": False}},
                "params": {n: {"prior": {"min": -1, "max": 1},
                               "ref": {"dist": "norm", "loc": 0, "scale": 0.001}} for n in names},
                "sampler": {"mcmc": {"covmat": np.diag(np.diag(cov)), "covmat_params": names,
                                     "measure_speeds": False, "learn_proposal": True,
                                     "burn_in": 0, "seed": 1, "max_samples": max_samples,
                                     "Rminus1_stop": 1e-9, "chains_per_gpu": chains}}}
        saved = sys.modules.get("cobaya.samplers.mcmc")
        plugin.install_as_mcmc()
        try:
            t0 = time.perf_counter()
            _, smp = run(info)
            wall = time.perf_counter() - t0
        finally:
            if saved is not None:
                sys.modules["cobaya.samplers.mcmc"] = saved
        props = int(smp.n_steps_raw) * chains
        rows = len(smp.collection)
        return {"value": props / wall, "unit": UNIT, "wall_s": wall, "proposals": props,
                "rows_in_collection": rows, "d2h_bytes": int(rows * smp._fm.row_width * 8),
                "what": f"cobaya.run.run(info), sampler: mcmc, chains_per_gpu={chains}, "
                        f"max_samples={max_samples}, no output prefix (wall clock of run())"}
    except Exception as e:  # never lose the benchmark line over the extra report
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        logging.disable(logging.NOTSET)


def run_gpu_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run "
                             "--nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as tdist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG / NCCL_DEBUG_FILE are left exactly as the launcher set them: whatever
        # NCCL logs (rank count, transports, NVLS) goes to stderr or to the launcher's file;
        # stdout stays the single JSON line (file descriptor 1 points at stderr, see
        # _quiet_stdout)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cobaya_b200 import problems
    from cobaya_b200.mcmc import EnsembleMCMC, TorchDist

    if world > 1:
        dist = TorchDist()
    prob = problems.get(args.config)
    fm, Dp = prob.fm, prob.fm.D
    C = args.chains
    x0 = prob.start(C, rank)
    locksteps = args.locksteps
    if locksteps is None:
        # about 10 ms of device work per step for every configuration
        locksteps = {"c1": 1024, "c2": 2 * fm.cycle_length, "c3": 20 * fm.cycle_length,
                     "c3x": 4 * fm.cycle_length}[prob.key]
    K, W = args.steps, args.warmup
    # stored rows per chain: acceptance stays below ~0.35 (and thinning only lowers it)
    rows_cap = int(0.45 * locksteps * (K + W + 2)) + 4096
    opts = {"seed": 1, "chains_per_gpu": C, "device": local, "rows_per_chain": rows_cap,
            "Rminus1_stop": 0.0, "learn_proposal_Rminus1_max": 1e9, "burn_in": 0}

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    smp = EnsembleMCMC(fm, x0, opts, dist=dist)
    eng = smp.engine
    n_ckpt = 0

    def one_step():
        nonlocal n_ckpt
        smp._ensure_row_capacity()
        eng.advance(locksteps)
        smp.n_steps_raw += locksteps
        g = smp._global_summary()
        smp._check_health(g)
        if smp.check_ready(g):
            smp.check_convergence_and_learn_proposal()
            smp.i_learn += 1
            n_ckpt += 1

    smp.launch_steps = locksteps
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    for _ in range(W):
        one_step()
    # the checkpoint kernels are part of the warm-up too (CUDA loads a kernel lazily on its
    # first launch; a first checkpoint inside the timed region would time the module load)
    smp._moments(0, 0)
    smp._bounds(0, 0, 0.475)
    barrier()
    if clocks:
        clocks.mark(0)
    eng.set_profiling(True)
    eng.kernel_times(reset=True)
    eng.window_counts(reset=True)
    l0 = eng.launch_count()
    rows0 = eng.summary()["sum_rows"]
    ck0 = n_ckpt
    eng.timer_start()
    t_wall = time.perf_counter()
    for _ in range(K):
        one_step()
    ms_dev = eng.timer_stop()
    barrier()
    if clocks:
        clocks.mark(1)
    t_wall = time.perf_counter() - t_wall
    kt = eng.kernel_times(reset=True)
    windows = eng.window_counts(reset=True)
    eng.set_profiling(False)
    launches = eng.launch_count() - l0
    rows1 = eng.summary()["sum_rows"]
    # max over ranks of the device time
    t = torch.tensor([ms_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_total = float(t.item())
    proposals = world * C * locksteps * K
    value = proposals / (ms_total * 1e-3)

    # ---- e2e: the same pass through the host API, HOST buffers on both sides.  Every step:
    # start points uploaded from page-locked host memory (set_state), `locksteps`
    # proposals per chain, then EVERY ROW the step stored is handed to the host -- device-side
    # compaction + one asynchronous D2H into a page-locked double buffer, running under the
    # next step's kernels (cb2_drain_start) -- plus the moments and the current points.
    # This is what the reference holds on the host after the same work (its
    # SampleCollection, collection.py:402-427), so nothing is left on the device.
    Wd = fm.row_width
    max_rows = int(0.5 * C * locksteps) + 1024
    bufs = [eng.host_buffer(max_rows * Wd) for _ in range(2)]
    cnts = [eng.host_buffer(C).view(np.int64) for _ in range(2)]
    x0_pinned = eng.host_buffer(C * Dp).reshape(C, Dp)
    x0_pinned[:] = x0
    Ke = max(3, min(K, 6))
    rows_out, sums, st = [], None, None
    for warm in (True, False):  # one untimed pass: staging buffers are sized, pages touched
        barrier()
        te = time.perf_counter()
        for k in range(1 if warm else Ke):
            eng.set_state(x0_pinned)                         # H2D: start points
            eng.advance(locksteps)
            if k >= 2:
                eng.drain_wait()                             # buffer k % 2 is free again
            n_out = eng.drain_start(bufs[k % 2], cnts[k % 2])   # D2H (async): the step's rows
            sums = smp._moments(0, 0)                        # moments (+ all-reduce) -> host
            st = eng.get_state()                             # D2H: current points + counters
            if not warm:
                rows_out.append(n_out)
        eng.drain_wait()
        barrier()
        te = time.perf_counter() - te
    tt = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    e2e_value = world * C * locksteps * Ke / float(tt.item())
    clk = clocks.stop() if clocks else None
    row_bytes = float(np.mean(rows_out)) * Wd * 8
    # the rows really are on the host: weights are positive integers, chain counts add up
    last = bufs[(Ke - 1) % 2][: rows_out[-1] * Wd].reshape(-1, Wd)
    e2e_ok = bool(rows_out[-1] == int(cnts[(Ke - 1) % 2].sum()) and np.all(last[:, 0] >= 1)
                  and np.all(last[:, 0] == np.round(last[:, 0])))
    h2d = x0_pinned.nbytes + Dp * 8
    d2h = (row_bytes + C * 8 + sums.nbytes + st["x"].nbytes + st["logpost"].nbytes
           + st["weight"].nbytes + st["n_rows"].nbytes + st["n_accepted"].nbytes
           + st["flags"].nbytes + 64)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (CUDA events bracketing every launch of the
    # class, on the engine's stream, inside the timed region)
    peak, peak_src = peaks()
    dom = max(("step", "basis"), key=lambda k: kt[k]["ms"])
    acc_rate = (rows1 - rows0) / (C * locksteps * K)     # stored-row rate a
    Wrow = fm.row_width
    bytes_per_prop = 16 * Dp + 8 * acc_rate * Wrow
    # the dominant kernel with ITS OWN bytes: the step kernel reads one basis column per
    # proposal and writes the stored rows; the basis kernel writes the bases (8 D B per
    # proposal, amortised) and reads the normals it is made from (about as much again)
    own_bytes = {"step": 8 * Dp + 8 * acc_rate * Wrow, "basis": 8 * Dp + 4 * (Dp + 2)}
    hot_ms = kt["step"]["ms"] + kt["basis"]["ms"]
    n_l = max(kt[dom]["launches"], 1)
    props_timed = C * locksteps * K
    props_per_launch = props_timed / max(kt["step"]["launches"], 1)
    achieved = bytes_per_prop * props_timed / (hot_ms * 1e-3) / 1e9
    dom_achieved = own_bytes[dom] * props_timed / (max(kt[dom]["ms"], 1e-9) * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and prob.key == "c1":
        try:
            tj = json.load(open(tp))
            # ncu figure of one window, scaled to the proposals of one launch now
            traffic = int(tj["dram_bytes_per_launch"] * props_per_launch /
                          float(tj.get("proposals_per_profiled_launch", props_per_launch)))
        except Exception:
            traffic = None
    flops = prob.flops_per_proposal * props_timed / (hot_ms * 1e-3) / 1e12
    fp64_pk, fp64_src = fp64_peak()
    if args.no_cpu_baseline:
        cb = None
    elif prob.key == "c3x":   # same function, built into the oracle
        p3 = problems.get("c3")
        cb = cpu_oracle_baseline(p3.fm, p3)
    else:
        cb = cpu_oracle_baseline(fm, prob)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": prob.workload, "D": Dp, "chains_per_gpu": C,
                   "locksteps_per_step": locksteps, "checkpoints_in_timed_region": n_ckpt - ck0,
                   "posterior_evaluations_per_proposal": prob.evals_per_proposal,
                   "l2": "inputs larger than L2: every window streams its Haar bases, their "
                         "normals and the new sample rows (> 1 GB at 8192 chains) through "
                         "the 126 MB L2; no explicit flush",
                   "step_kernel": ("dmma-dragging (k_step_drag)"
                                   if fm.drag and eng.last_step_kernel() == 1
                                   else KERNEL_NAMES.get(eng.last_step_kernel(), "?")),
                   "windows_by_step_kernel": windows,
                   "parallelism": f"chains sharded over {world} GPU(s), no data-path "
                                  "collective; NCCL all-reduce of moments per checkpoint"},
        "clocks": clk, "gpu_launches": int(launches), "engine_note": eng.debug_message(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": Ke,
                "rows_per_step": float(np.mean(rows_out)), "rows_checked_on_host": e2e_ok,
                "d2h_gbs": d2h * Ke / float(tt.item()) / 1e9,
                "what": "set_state(page-locked host) + advance + EVERY stored row drained to "
                        "page-locked host memory (async, under the next step) + "
                        "moments->host + get_state; bound by the host link, not the kernels"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "dominant_kernel": dom,
                     "binding_roof": "fp64 tensor pipe (DMMA m8n8k4; see roofline.fp64.frac), "
                                     "not HBM: the HBM figures follow the contract's recipe",
                     "dominant_kernel_own": {"bytes_per_proposal": own_bytes[dom],
                                             "achieved": dom_achieved,
                                             "frac": dom_achieved / peak},
                     "algorithmic_bytes_per_proposal": bytes_per_prop,
                     "stored_row_rate": acc_rate,
                     "avg_launch_ms": {k: kt[k]["ms"] / max(kt[k]["launches"], 1)
                                       for k in kt},
                     "share_of_step": {k: kt[k]["ms"] / ms_dev for k in kt},
                     "proposals_per_step_launch": props_per_launch, "launches": n_l,
                     "fp64": {"tflops": flops, "peak_tflops": fp64_pk,
                              "peak_source": fp64_src, "frac": flops / fp64_pk,
                              "flops_per_proposal": prob.flops_per_proposal,
                              "note": "binding roof above D~23 (SURVEY.md 8d); DMMA and "
                                      "vector FP64 share one datapath"}},
        "cpu_baseline": cb,
        "wall_s_timed_region": t_wall,
    }
    if prob.evals_per_proposal > 1:
        line["posterior_evaluations_per_s"] = value * prob.evals_per_proposal
    try:  # R-1 of means at the convergence checks of this run (warm-up + timed region)
        def _fin(v):
            return float(v) if v is not None and np.isfinite(v) else None

        line["convergence"] = [
            {"accepted_steps": int(c.N), "Rminus1": _fin(c.Rminus1),
             "acceptance_rate": _fin(c.acceptance_rate), "covmat_learned": bool(c.learned)}
            for c in smp.progress]
    except Exception:  # never lose the benchmark line over the extra report
        pass
    if world == 1 and prob.key == "c1" and not args.no_cobaya_run:
        eng.close()
        line["e2e_cobaya_run"] = e2e_through_cobaya_run(C)
    emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torch warnings) write to file descriptor 1 from
    native code: point fd 1 at stderr for the whole run and keep the original for the one
    JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--locksteps", type=int, default=None)
    ap.add_argument("--config", default="c1", choices=["c1", "c2", "c3", "c3x"],
                    help="c1 = BASELINE configs[1] (headline), c2 = configs[2] (128-D, 3 modes, "
                         "speed blocks), c3 = configs[3] (30-D Rosenbrock, dragging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cobaya-run", action="store_true",
                    help="skip the second end-to-end figure through cobaya.run.run")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
        try:  # torchrun ranks: leave NCCL cleanly
            import torch.distributed as tdist

            if tdist.is_available() and tdist.is_initialized():
                tdist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
