#!/usr/bin/env python
"""BASELINE.json configs[4]: D in {8, 32, 128, 512} x chains in {1k, 8k, 64k} (plus the
headline D = 64), single-mode correlated Gaussian, one block, proposals/s against the HBM
and FP64 roofs of SURVEY.md section 8d.  One JSON line per cell; device time from CUDA events
on the engine's stream.

    python tools/dim_sweep.py [--cells 8x1024,64x8192,...] [--cycles 4] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cobaya_b200.engine import Engine
from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

KERNEL = {0: "general", 1: "dmma", 2: "dmma-producer-consumer", 3: "dmma-streamed"}


def peaks():
    hbm, fp64 = 6552.0, 36.9
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = float(json.load(open(p)).get("hbm_gbs", hbm))
    return hbm, fp64  # fp64: tools/fp64_peak.cu on this pool's B200 (profiles/r1_fp64_peak.json)


def run_cell(D, C, cycles, seed=1):
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
    rng = np.random.default_rng(0)
    L = np.linalg.cholesky(cov)
    x0 = rng.standard_normal((C, D)) @ L.T
    n = cycles * D
    warm = n  # same call shape as the timed one: every window buffer is sized before timing
    if D >= 512 and C > 8192:
        warm = 64  # 64k chains at D = 512: windows are chunked over the chains; keep the rows small
    # every chain keeps every row it stores (the stored-row rate stays below 0.35): a cell
    # whose chains ran out of room would have done less work than it claims
    cap = int(0.45 * (n + warm)) + 64
    if C * cap * (D + 6) * 8 > 120e9:
        raise RuntimeError(f"row store of {C * cap * (D + 6) * 8 / 1e9:.0f} GB: lower --cycles")
    eng = Engine(fm, n_chains=C, seed=seed, rows_cap=cap)
    eng.set_state(x0)
    eng.advance(warm)
    eng.sync()
    eng.timer_start()
    eng.advance(n)
    ms = eng.timer_stop()
    s = eng.summary()
    st_rate = s["sum_rows"] / float(C * (n + warm))
    rate = C * n / (ms * 1e-3)
    hbm, fp64 = peaks()
    bytes_pp = 16.0 * D + 8.0 * st_rate * (D + 6)
    out = {"D": D, "chains": C, "proposals": C * n, "ms": ms, "proposals_per_s": rate,
           "step_kernel": KERNEL.get(eng.last_step_kernel(), "?"),
           "stored_row_rate": st_rate,
           "hbm_frac": bytes_pp * rate / (hbm * 1e9),
           "fp64_frac": 4.0 * D * D * rate / (fp64 * 1e12),
           "n_stuck": s["n_stuck"], "n_rows_full": s["n_rows_full"],
           "windows_by_step_kernel": eng.window_counts(), "engine_note": eng.debug_message()}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", default="")
    ap.add_argument("--cycles", type=int, default=4)
    a = ap.parse_args()
    if a.cells:
        cells = [tuple(int(v) for v in c.split("x")) for c in a.cells.split(",")]
    else:
        cells = [(D, C) for D in (8, 32, 64, 128, 512) for C in (1024, 8192, 65536)]
    for D, C in cells:
        cyc = a.cycles if D < 512 else max(1, a.cycles // 2)
        if D >= 512 and C > 8192:
            cyc = 1  # 137 GB of bases for 64k chains: cb2_advance chunks the window over the chains
        try:
            print(json.dumps(run_cell(D, C, cyc)), flush=True)
        except Exception as e:  # report the cell, go on with the grid
            print(json.dumps({"D": D, "chains": C, "error": str(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
