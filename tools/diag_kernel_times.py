import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np
from cobaya_b200.engine import Engine
from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
for D, C in ((128, 8192), (512, 1024)):
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=cov)
    x0 = np.random.default_rng(0).standard_normal((C, D)) @ np.linalg.cholesky(cov).T
    eng = Engine(fm, n_chains=C, seed=1, rows_cap=2 * D)
    eng.set_state(x0); eng.advance(D); eng.sync()
    eng.set_profiling(True); eng.kernel_times(reset=True)
    eng.advance(2 * D); eng.sync()
    print(D, C, json.dumps(eng.kernel_times()))
    eng.close()
