"""Diagnostic: time the checkpoint kernels in isolation after a realistic run."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cobaya_b200.engine import Engine
from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

D, C = 64, 8192
cov = synthetic_gaussian_cov(D)
fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
x0 = np.random.default_rng(0).multivariate_normal(np.zeros(D), cov, size=C)
eng = Engine(fm, n_chains=C, seed=1, rows_cap=6000)
eng.set_state(x0)
for steps in (8192, 8192):
    eng.timer_start(); eng.advance(steps); ms = eng.timer_stop()
    print(f"advance {steps}: {ms:.1f} ms -> {C*steps/ms/1e3:.3e} proposals/s", eng.summary())
    for rep in range(3):
        eng.set_profiling(True); eng.kernel_times(reset=True)
        eng.timer_start(); s = eng.moments(); ms = eng.timer_stop()
        kt = eng.kernel_times(reset=True); eng.set_profiling(False)
        print(f"  moments rep {rep}: timer {ms:.2f} ms, bracket {kt['moments']}")
    t = time.perf_counter(); b = eng.bounds(0.475); dt = time.perf_counter() - t
    print(f"  bounds: wall {dt*1e3:.1f} ms")
for pol, name in ((0, "wy basis"), (4, "dfma basis")):
    eng.set_kernel_policy(pol)
    eng.set_profiling(True); eng.kernel_times(reset=True)
    eng.advance(1024); kt = eng.kernel_times(reset=True); eng.set_profiling(False)
    print(name, {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in kt.items()})
