#!/usr/bin/env python
"""Throughput of the external-function route (propose / user kernel / accept launches per
proposal) next to the built-in kernels on the same D-dimensional Gaussian, 8192 chains."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cobaya_b200.engine import Engine
from tests import ext_functions


def rate(fm, x0, policy, n=512):
    e = Engine(fm, n_chains=len(x0), seed=1, chain_id0=0, rows_want=600)
    if policy is not None:
        e.set_kernel_policy(policy)
    e.set_state(x0)
    e.advance(64)
    e.sync()
    t0 = time.perf_counter()
    e.advance(n)
    e.sync()
    dt = time.perf_counter() - t0
    e.close()
    return len(x0) * n / dt


def main():
    C = 8192
    out = []
    for D in (8, 32, 64):
        builtin, ext, mu, cov = ext_functions.gaussian_pair(D)
        x0 = np.random.default_rng(1).multivariate_normal(mu, cov, size=C)
        out.append({"D": D, "chains": C,
                    "external_function_route": rate(ext, x0, None),
                    "builtin_general_kernel": rate(builtin, x0, 1),
                    "builtin_default_kernel": rate(builtin, x0, None)})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
