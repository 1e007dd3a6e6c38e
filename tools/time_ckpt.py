#!/usr/bin/env python
"""Time the checkpoint of the benchmark shape (64-D, 8192 chains): SYRK (cb2_moments), the
device algebra (cb2_checkpoint_device), the device repack (cb2_adopt_proposal) and the host
route they replace (sums -> host, LAPACK, cb2_set_proposal + rebuild)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobaya_b200 import problems
from cobaya_b200.convergence import rminus1_from_sums
from cobaya_b200.engine import Engine


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    p = problems.config1(D)
    eng = Engine(p.fm, n_chains=C, seed=1, chain_id0=0, rows_want=1500)
    eng.set_state(p.start(C, 0))
    eng.advance(2048)
    eng.sync()
    shift = np.zeros(D)
    out = {"D": D, "chains": C, "rows_per_chain_mean": float(eng.get_state()["n_rows"].mean())}

    def t(fn, n=5):
        fn()
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        eng.sync()
        return (time.perf_counter() - t0) / n * 1e3

    out["moments_device_ms"] = t(lambda: eng.moments(shift=shift, host=False))
    out["moments_to_host_ms"] = t(lambda: eng.moments(shift=shift, host=True))
    eng.moments(shift=shift, host=False)
    out["checkpoint_device_ms"] = t(lambda: eng.checkpoint_device())
    res = eng.checkpoint_device()
    out["jacobi_sweeps"] = res["sweeps"]

    def adopt():
        eng.checkpoint_device()
        eng.adopt_proposal()
    out["checkpoint_device_plus_adopt_ms"] = t(adopt)
    sums = eng.moments(shift=shift, host=True)

    def host_route():
        s = eng.moments(shift=shift, host=True)
        r = rminus1_from_sums(s, D, shift)
        eng.set_covariance(r["W"])
        eng.advance(0) if False else None
    out["host_route_incl_moments_ms"] = t(host_route)
    out["host_algebra_only_ms"] = t(lambda: rminus1_from_sums(sums, D, shift))
    r = rminus1_from_sums(sums, D, shift)
    out["host_set_covariance_ms"] = t(lambda: eng.set_covariance(r["W"]))
    # the rebuild of the packs happens at the next advance
    def adv():
        eng.set_covariance(r["W"])
        eng.advance(1)
    out["host_set_covariance_plus_one_proposal_ms"] = t(adv)
    out["one_proposal_ms"] = t(lambda: eng.advance(1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
