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


PLUGIN_WORKER = textwrap.dedent("""
    import os, sys, json, copy
    import numpy as np
    sys.path.insert(0, {root!r})
    for p in ({root!r} + "/oracle/shims", {root!r} + "/baseline/_ref"):
        sys.path.insert(0, p)
    import cobaya_b200.distributed as cbd
    cbd.init()                                   # NCCL process group + cobaya.mpi on top of it
    from cobaya import mpi
    from cobaya.run import run
    from tests.util import load_golden
    rank = mpi.rank()
    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)

    def info(prefix, max_samples):
        return {{"likelihood": {{"gaussian_mixture": {{
                    "means": [mean], "covs": [cov], "input_params_prefix": "a_",
                    "output_params_prefix": "", "derived": True}}}},
                "params": dict({{f"a__{{i}}": {{"prior": {{"min": -1, "max": 1}}}} for i in range(3)}},
                               **{{f"_{{i}}": None for i in range(3)}}),
                "sampler": {{"cobaya_b200.plugin.MCMC": {{
                    "covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
                    "burn_in": 5, "max_tries": 3000, "learn_proposal_Rminus1_max": 30,
                    "Rminus1_stop": 1e-9, "measure_speeds": False, "seed": 5,
                    "chains_per_gpu": 16, "max_samples": max_samples}}}}, "output": prefix}}

    pb = os.path.join({outdir!r}, "b", "run")
    _, first = run(copy.deepcopy(info(pb, 400)), force=True)
    n_first = len(first.collection)
    _, second = run(copy.deepcopy(info(pb, 900)), resume=True)
    pa = os.path.join({outdir!r}, "a", "run")
    _, full = run(copy.deepcopy(info(pa, 900)), force=True)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    same = bool(np.array_equal(key(full.collection.data.to_numpy()),
                               key(second.collection.data.to_numpy())))
    with open(os.path.join({outdir!r}, f"res{{rank}}.json"), "w") as f:
        json.dump(dict(rank=rank, name=first.collection.name, n_first=n_first, same=same,
                       n_second=len(second.collection), n_full=len(full.collection),
                       R=[float(v) for v in second.progress["Rminus1"]],
                       device=int(second._ens.engine.device)), f)
""")


def test_two_gpus_plugin_through_cobaya_run_with_output_and_resume(tmp_path):
    """The multi-process boundary on real GPUs: torchrun x 2, NCCL data plane, cobaya.mpi on
    torch.distributed (cobaya_b200.distributed), per-rank chain files, root-only checkpoint
    with mpi_size 2, bit-exact resume on both ranks (mcmc.py:131-151,1045-1078)."""
    import torch

    from tests.refenv import enable_reference

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    enable_reference()
    script = tmp_path / "w.py"
    script.write_text(PLUGIN_WORKER.format(root=ROOT, outdir=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    res = [json.load(open(tmp_path / f"res{r}.json")) for r in range(2)]
    assert [r["name"] for r in res] == ["1", "2"] and [r["device"] for r in res] == [0, 1]
    files = sorted(os.listdir(tmp_path / "b"))
    for want in ["run.1.txt", "run.2.txt", "run.checkpoint", "run.covmat", "run.progress",
                 "run.b200_state.1.npz", "run.b200_state.2.npz"]:
        assert want in files, (want, files)
    import yaml

    ck = yaml.safe_load(open(tmp_path / "b" / "run.checkpoint"))
    assert ck["sampler"]["cobaya_b200.plugin.MCMC"]["mpi_size"] == 2
    for r in res:
        assert r["same"] and r["n_second"] == r["n_full"] > r["n_first"] > 0
    assert res[0]["R"] == res[1]["R"]  # one all-reduce, identical verdicts
