#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the kernels added in round 2 that the ncu/bench
# runs do not otherwise exercise under a checker: small instances of the device checkpoint, the
# external-function route and the dragging kernel.  Usage: bash tools/sanitize_new_kernels.sh
out=gpurun_out
mkdir -p $out
cat > $out/_san_driver.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from cobaya_b200 import problems
from cobaya_b200.engine import Engine
from tests import ext_functions
which = sys.argv[1]
if which == "ckpt":
    for D in (13, 32):
        p = problems.config1(D)
        e = Engine(p.fm, n_chains=64, seed=3, chain_id0=0, rows_cap=512)
        e.set_state(p.start(64, 0)); e.advance(200)
        e.moments(shift=np.zeros(D), host=False)
        r = e.checkpoint_device(); e.adopt_proposal(); e.advance(40)
        print("ckpt", D, r["Rminus1"], r["sweeps"], e.get_state()["flags"].any())
elif which == "ext":
    b, x, mu, cov = ext_functions.gaussian_pair(8)
    e = Engine(x, n_chains=40, seed=3, chain_id0=0, rows_cap=512)
    e.set_state(np.random.default_rng(1).multivariate_normal(mu, cov, size=40)); e.advance(60)
    print("ext", e.get_state()["n_rows"].sum(), e.get_state()["flags"].any())
elif which == "drag":
    p = problems.config3()
    e = Engine(p.fm, n_chains=40, seed=3, chain_id0=0, rows_cap=512)
    e.set_state(p.start(40, 0)); e.advance(30)
    print("drag", e.last_step_kernel(), e.get_state()["n_rows"].sum(), e.get_state()["flags"].any())
PY
for tool in memcheck racecheck; do
  for which in ckpt ext drag; do
    timeout -s KILL 280 compute-sanitizer --tool $tool --print-limit 5 python $out/_san_driver.py $which \
        > $out/sanitizer_${tool}_${which}.txt 2>&1
    echo "== $tool $which: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitizer_${tool}_${which}.txt | tail -1)"
    grep -E "^(ckpt|ext|drag) " $out/sanitizer_${tool}_${which}.txt | tail -2
  done
done
