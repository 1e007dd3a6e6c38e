"""Checkpoint cost at D = 64 / 128 / 256: time cb2_moments over ~n rows per chain."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cobaya_b200.engine import Engine
from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
for D, C, steps in ((64, 8192, 4096), (128, 8192, 4096), (256, 2048, 4096)):
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
    x0 = np.random.default_rng(0).standard_normal((C, D)) @ np.linalg.cholesky(cov).T
    eng = Engine(fm, n_chains=C, seed=1, rows_cap=int(0.45 * steps) + 64)
    eng.set_state(x0); eng.advance(steps); eng.sync()
    rows = eng.summary()["sum_rows"] / C
    eng.moments(shift=np.zeros(D))
    t = time.perf_counter(); eng.moments(shift=np.zeros(D)); dt = time.perf_counter() - t
    fl = C * (rows / 2) * D * D * 2
    print(json.dumps({"D": D, "chains": C, "rows_per_chain": rows, "moments_ms": dt * 1e3,
                      "tflops_full_matrix": fl / dt / 1e12}))
    eng.close()
