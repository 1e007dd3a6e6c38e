#!/usr/bin/env python
"""R-1 history of the UNMODIFIED reference (cobaya 3.6.2 from baseline/_ref, one chain, single-chain
split rule mcmc.py:795-822) on the benchmark target (64-D correlated Gaussian, diagonal start
covmat, learning on), for the "R-1 convergence vs ref" half of BASELINE.json's metric.  CPU only.

    python tools/reference_convergence.py [max_samples] > profiles/r1_reference_convergence.json
"""
import json
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
os.environ["COBAYA_NOMPI"] = "1"
logging.disable(logging.CRITICAL)
import numpy as np  # noqa: E402

from cobaya.model import get_model  # noqa: E402
from cobaya.sampler import get_sampler  # noqa: E402

from cobaya_b200.flatmodel import synthetic_gaussian_cov  # noqa: E402

D = 64
n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
cov = synthetic_gaussian_cov(D)
names = [f"x{i}" for i in range(D)]
info = {"likelihood": {"gaussian_mixture": {"means": [np.zeros(D)], "covs": [cov],
                                            "input_params": names, "derived": False}},
        "params": {n: {"prior": {"min": -1, "max": 1},
                       "ref": {"dist": "norm", "loc": 0, "scale": 0.001}} for n in names},
        "sampler": {"mcmc": {"covmat": np.diag(np.diag(cov)), "covmat_params": names,
                             "measure_speeds": False, "learn_proposal": True, "burn_in": 0,
                             "seed": 1, "max_samples": n_samples, "Rminus1_stop": 1e-9,
                             "output_every": "1000s"}}}
model = get_model(info)
sampler = get_sampler(info["sampler"], model)
t = time.perf_counter()
sampler.run()
dt = time.perf_counter() - t
prog = sampler.products()["progress"]
col = sampler.products()["sample"]
half = col.skip_samples(0.5, inplace=False) if hasattr(col, "skip_samples") else col
m, S = half.mean(), half.cov()
sig = np.sqrt(np.diag(cov))
out = {
    "what": "unmodified reference, 1 chain, 64-D benchmark target, R-1 by the single-chain split rule",
    "proposals": int(sampler.n_steps_raw), "accepted_steps": int(len(col)), "seconds": dt,
    "proposals_per_s": sampler.n_steps_raw / dt,
    "checkpoints": [{"accepted_steps": int(r.N), "Rminus1": None if not np.isfinite(float(r.Rminus1)) else float(r.Rminus1),
                     "acceptance_rate": float(r.acceptance_rate)} for r in prog.itertuples()],
    "final_max_abs_mean_error_in_sigma": float(np.max(np.abs(np.asarray(m)) / sig)),
    "final_max_rel_variance_error": float(np.max(np.abs(np.diag(np.asarray(S)) / np.diag(cov) - 1))),
}
print(json.dumps(out, indent=1))
