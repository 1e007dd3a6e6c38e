#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[3] at full size on one GPU (8192 chains): proposals/s
(and posterior evaluations/s for dragging).  Device time from CUDA events.

    python tools/configs_bench.py > gpurun_out/configs.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cobaya_b200.engine import Engine
from cobaya_b200.flatmodel import FlatModel, LikeSpec

KERNEL = {0: "general", 1: "dmma", 2: "dmma-producer-consumer", 3: "dmma-streamed"}


def mixture_cov(D, rng, scale=0.02):
    A = rng.standard_normal((D, 2 * D))
    C = A @ A.T / (2 * D)
    d = np.sqrt(np.diag(C))
    s = scale * 10 ** rng.uniform(-0.5, 0.5, D)
    return (C / d[:, None] / d[None, :]) * s[:, None] * s[None, :]


def config3():
    rng = np.random.default_rng(20260925)
    D, n_slow = 128, 32
    a = LikeSpec.gaussian_mixture(np.arange(n_slow),
                                  [np.full(n_slow, 0.03 * k) for k in range(3)],
                                  [mixture_cov(n_slow, rng) for _ in range(3)], name="slow")
    b = LikeSpec.gaussian_mixture(np.arange(n_slow, D),
                                  [np.full(D - n_slow, 0.03 * k) for k in range(3)],
                                  [mixture_cov(D - n_slow, rng) for _ in range(3)], name="fast")
    blocks = [list(range(n_slow)), list(range(n_slow, D))]
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -1.0), upper=np.full(D, 1.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[a, b],
                   blocks=blocks, oversampling=[1, 3],
                   proposal_cov=np.diag(np.full(D, 0.01**2)), output_thin=2)
    return "configs[2]: 128-D, two 3-mode gaussian_mixture components, blocks 32+96, oversampling [1,3]", fm, 0.01, 1


def config4():
    D, n_slow, o_fast = 30, 10, 4
    lk = LikeSpec.rosenbrock(np.arange(D), scale=1.0 / 20.0)
    n_drag = int(np.round(o_fast * (D - n_slow) / n_slow))
    fm = FlatModel(names=[f"x{i}" for i in range(D)], prior_kind=np.zeros(D, np.int32),
                   lower=np.full(D, -5.0), upper=np.full(D, 5.0), loc=np.zeros(D),
                   pscale=np.ones(D), periodic=np.zeros(D, np.int32), likes=[lk],
                   blocks=[list(range(n_slow)), list(range(n_slow, D))],
                   oversampling=[1, o_fast], drag=True, i_last_slow_block=0,
                   drag_interp_steps=n_drag, proposal_cov=np.diag(np.full(D, 0.05**2)))
    return "configs[3]: 30-D Rosenbrock, dragging (n_drag = %d)" % n_drag, fm, 0.05, 2 * n_drag + 1


def main():
    C = 8192
    for name, fm, spread, evals_per_prop in (config3(), config4()):
        D = fm.D
        rng = np.random.default_rng(1)
        x0 = (1.0 if fm.drag else 0.0) + rng.normal(0, spread, (C, D))
        n = 2 * fm.cycle_length if not fm.drag else 200
        eng = Engine(fm, n_chains=C, seed=1, rows_cap=n + 64)
        eng.set_state(x0)
        eng.advance(n // 2)
        eng.sync()
        eng.timer_start()
        eng.advance(n)
        ms = eng.timer_stop()
        s = eng.summary()
        rate = C * n / (ms * 1e-3)
        print(json.dumps({"config": name, "D": D, "chains": C, "proposals": C * n, "ms": ms,
                          "proposals_per_s": rate,
                          "posterior_evals_per_s": rate * evals_per_prop,
                          "step_kernel": KERNEL.get(eng.last_step_kernel(), "?"),
                          "acceptance": s["sum_accepted"] / float(C * (n + n // 2)),
                          "n_stuck": s["n_stuck"]}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
