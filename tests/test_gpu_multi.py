"""
Multi-GPU test (needs >= 2 GPUs on the box; skipped otherwise): two ranks under torchrun,
NCCL all-reduce of the per-GPU checkpoint statistics.  Every rank must obtain bit-identical
R-1 / W, equal to what ONE engine holding all the chains computes (chains are sharded by
global id, so the union of the two ranks' chains is the same ensemble).
"""

import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tests.util import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cobaya_b200.mcmc import EnsembleMCMC, TorchDist
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    D, C = 16, 64
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=np.diag(np.diag(cov)))
    x0_all = np.random.default_rng(0).multivariate_normal(np.zeros(D), cov, size=2 * C)
    td = TorchDist()
    s = EnsembleMCMC(fm, x0_all[td.rank * C:(td.rank + 1) * C],
                     dict(seed=5, chains_per_gpu=C, device=local, rows_per_chain=3000,
                          max_samples=800, Rminus1_stop=1e-9), dist=td)
    s.run()
    out = dict(rank=td.rank, n_chains=s.n_chains, R=[c.Rminus1 for c in s.progress],
               N=[c.N for c in s.progress], learned=[c.learned for c in s.progress],
               cov00=float(s.fm.get_covariance()[0, 0]), steps=s.n_steps_raw)
    with open(os.path.join({outdir!r}, f"rank{{td.rank}}.json"), "w") as f:
        json.dump(out, f)
    dist.destroy_process_group()
""")


def _free_port():
    import socket

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def test_two_gpus_equal_one_engine_with_all_chains(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, outdir=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    assert len(res) == 2 and res[0]["n_chains"] == 128
    assert res[0]["R"] == res[1]["R"] and res[0]["cov00"] == res[1]["cov00"]
    assert res[0]["N"] == res[1]["N"] and len(res[0]["R"]) >= 1
    # one engine with all 128 chains (same global ids, same seed) -> same ensemble
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    from cobaya_b200.mcmc import EnsembleMCMC

    D, C = 16, 64
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.zeros(D), cov, proposal_cov=np.diag(np.diag(cov)))
    x0_all = np.random.default_rng(0).multivariate_normal(np.zeros(D), cov, size=2 * C)
    s = EnsembleMCMC(fm, x0_all, dict(seed=5, chains_per_gpu=2 * C, rows_per_chain=3000,
                                      max_samples=800, Rminus1_stop=1e-9)).run()
    assert [c.N for c in s.progress] == res[0]["N"]
    np.testing.assert_allclose([c.Rminus1 for c in s.progress], res[0]["R"], rtol=1e-8)
    np.testing.assert_allclose(s.fm.get_covariance()[0, 0], res[0]["cov00"], rtol=1e-10)
