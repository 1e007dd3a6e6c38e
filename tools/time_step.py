#!/usr/bin/env python
"""Per-kernel-class device time of the headline window (64-D, 8192 chains, 256 proposals per
launch) for the library selected by COBAYA_B200_LIB -- kernel experiments.
    python tools/time_step.py [policy] [windows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cobaya_b200.engine import Engine
from cobaya_b200 import problems

policy = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prob = problems.get("c1")
C = 8192
eng = Engine(prob.fm, n_chains=C, seed=1, rows_cap=256 * (nwin + 4) + 8)
eng.set_kernel_policy(policy)
eng.set_state(prob.start(C, 0))
eng.advance(512)
eng.sync()
import ctypes
dbg = np.zeros(16, np.int64)
eng.lib.cb2_debug_counters(eng.h, dbg.ctypes.data_as(ctypes.c_void_p), 1)
eng.set_profiling(True)
eng.kernel_times(reset=True)
eng.timer_start()
eng.advance(256 * nwin)
ms = eng.timer_stop()
kt = eng.kernel_times()
print(os.path.basename(os.environ.get("COBAYA_B200_LIB", "default")), "policy", policy,
      "total ms/window %.4f" % (ms / nwin),
      " ".join("%s %.4f" % (k, v["ms"] / max(v["launches"], 1)) for k, v in kt.items()),
      "prop/s %.4g" % (C * 256 * nwin / ms * 1e3), eng.window_counts())
eng.lib.cb2_debug_counters(eng.h, dbg.ctypes.data_as(ctypes.c_void_p), 0)
if dbg.any():
    per = 256.0 * nwin
    print("consumer cycles/step: wait %.0f bulk %.0f tail %.0f store+release %.0f | producer cycles/item: "
          "waits %.0f mma %.0f oempty %.0f store %.0f" % (*(dbg[0:4] / per), *(dbg[8:12] / (per * 7 / 2.0))))
