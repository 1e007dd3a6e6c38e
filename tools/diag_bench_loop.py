"""Diagnostic: replicate bench.py's loop and time the checkpoint pieces with wall clock."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cobaya_b200.mcmc import EnsembleMCMC

fm, cov = bench.build_problem()
C, locksteps, K, W = 8192, 1024, 20, 3
x0 = bench.start_points(fm, cov, C, 0)
rows_cap = int(0.45 * locksteps * (K + W + 2)) + 4096
pol = int(sys.argv[1]) if len(sys.argv) > 1 else 0
smp = EnsembleMCMC(fm, x0, {"seed": 1, "chains_per_gpu": C, "device": 0,
                            "rows_per_chain": rows_cap, "Rminus1_stop": 0.0,
                            "learn_proposal_Rminus1_max": 1e9, "burn_in": 0})
eng = smp.engine
eng.set_kernel_policy(pol)
for step in range(K + W):
    t0 = time.perf_counter()
    eng.advance(locksteps); eng.sync()
    t1 = time.perf_counter()
    g = smp._global_summary()
    if smp.check_ready(g):
        ta = time.perf_counter()
        sums = eng.moments(shift=smp._shift)
        tb = time.perf_counter()
        sums2 = eng.moments(shift=smp._shift)
        tc = time.perf_counter()
        smp.check_convergence_and_learn_proposal()
        td = time.perf_counter()
        smp.i_learn += 1
        print(f"step {step}: advance {1e3*(t1-t0):.1f} ms; moments#1 {1e3*(tb-ta):.1f} ms, "
              f"moments#2 {1e3*(tc-tb):.1f} ms, full checkpoint {1e3*(td-tc):.1f} ms; "
              f"min_rows {g['min_rows']} kernel {eng.last_step_kernel()}")
    elif step % 5 == 0:
        print(f"step {step}: advance {1e3*(t1-t0):.1f} ms")
