"""
CPU tests (gloo, world_size 2) of the multi-process boundary: the plugin driven through
``cobaya.run.run`` with an ``output`` prefix and then resumed, one process per (absent)
GPU.  The chains are advanced by the oracle-backed test engine (tests/oracle_engine.py);
everything else is the product's host code: ``cobaya_b200.distributed`` (cobaya.mpi on
torch.distributed), per-rank chain files, root-only ``.checkpoint/.covmat/.progress``,
``mpi_size`` of the checkpoint, per-rank start-point RNG, growth of the row store, and the
"a failing rank ends the others" rule.  References: cobaya/mpi.py:231-267,350-467,
cobaya/samplers/mcmc/mcmc.py:131-151,1045-1078, cobaya/sampler.py:369-384.
"""

import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tests.refenv import have_reference
from tests.util import ROOT

pytestmark = pytest.mark.skipif(not have_reference(),
                                reason="reference package not installed under baseline/_ref")


def _free_port():
    import socket

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _torchrun(script, n=2, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    return subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        capture_output=True, text=True, env=env, timeout=timeout)


PRELUDE = textwrap.dedent("""
    import os, sys, json, copy
    import numpy as np
    sys.path.insert(0, {root!r})
    for p in ({root!r} + "/oracle/shims", {root!r} + "/baseline/_ref"):
        sys.path.insert(0, p)
    import logging
    import cobaya_b200.distributed as cbd
    comm = cbd.init(backend="gloo")
    from cobaya import mpi
    from cobaya.run import run
    import cobaya_b200.plugin as plugin
    from tests.oracle_engine import OracleEngine
    from tests.util import load_golden
    plugin.MCMC._engine_factory = OracleEngine     # no GPU in this container
    rank = mpi.rank()
    out = {outdir!r}
    g = load_golden("g1_gauss3d")
    mean, cov = g["means"][0], np.asarray(g["covs"]).reshape(3, 3)

    def info(prefix, max_samples, **extra):
        opts = {{"covmat": np.asarray(g["S0"]), "covmat_params": ["a__0", "a__1", "a__2"],
                "burn_in": 2, "max_tries": 3000, "learn_proposal_Rminus1_max": 30,
                "Rminus1_stop": 1e-9, "measure_speeds": False, "seed": 5,
                "chains_per_gpu": 3, "max_samples": max_samples, "learn_every": "20d",
                "rows_per_chain": 40}}
        opts.update(extra)
        return {{"likelihood": {{"gaussian_mixture": {{
                    "means": [mean], "covs": [cov], "input_params_prefix": "a_",
                    "output_params_prefix": "", "derived": True}}}},
                "params": dict({{f"a__{{i}}": {{"prior": {{"min": -1, "max": 1}}}} for i in range(3)}},
                               **{{f"_{{i}}": None for i in range(3)}}),
                "sampler": {{"cobaya_b200.plugin.MCMC": opts}}, "output": prefix}}
""")


RUN_AND_RESUME = PRELUDE + textwrap.dedent("""
    assert mpi.size() == 2 and mpi.is_main_process() == (rank == 0)
    pb = os.path.join(out, "b", "run")
    _, first = run(copy.deepcopy(info(pb, 150)), force=True)
    n_first = len(first.collection)
    x0 = first._x0.copy()
    grown = first._ens.engine.grown
    _, second = run(copy.deepcopy(info(pb, 300)), resume=True)
    pa = os.path.join(out, "a", "run")
    _, full = run(copy.deepcopy(info(pa, 300)), force=True)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    same = bool(np.array_equal(key(full.collection.data.to_numpy()),
                               key(second.collection.data.to_numpy())))
    res = dict(rank=rank, name=first.collection.name, n_first=n_first,
               n_second=len(second.collection), n_full=len(full.collection), same=same,
               x0=x0.tolist(), grown=int(grown), steps=[second.n_steps_raw, full.n_steps_raw],
               cov_equal=bool(np.array_equal(second.proposer.get_covariance(),
                                             full.proposer.get_covariance())),
               n_progress=[len(second.progress), len(full.progress)])
    with open(os.path.join(out, f"res{{rank}}.json"), "w") as f:
        json.dump(res, f)
""")


def test_two_ranks_through_cobaya_run_with_output_and_resume(tmp_path):
    from tests.refenv import enable_reference

    enable_reference()
    script = tmp_path / "w.py"
    script.write_text(RUN_AND_RESUME.format(root=ROOT, outdir=str(tmp_path)))
    p = _torchrun(script)
    assert p.returncode == 0, p.stderr[-4000:]
    res = [json.load(open(tmp_path / f"res{r}.json")) for r in range(2)]
    # one chain file per process, named by rank (mcmc.py:142), nothing overwritten
    assert [r["name"] for r in res] == ["1", "2"]
    files = sorted(os.listdir(tmp_path / "b"))
    for want in ["run.1.txt", "run.2.txt", "run.checkpoint", "run.covmat", "run.progress",
                 "run.b200_state.1.npz", "run.b200_state.2.npz", "run.b200_rows.1.bin",
                 "run.b200_rows.2.bin", "run.updated.yaml"]:
        assert want in files, (want, files)
    # root-only checkpoint with the number of processes (mcmc.py:131-139,1045-1078)
    import yaml

    ck = yaml.safe_load(open(tmp_path / "b" / "run.checkpoint"))
    assert ck["sampler"]["cobaya_b200.plugin.MCMC"]["mpi_size"] == 2
    # distinct start points per process (one SeedSequence child each, sampler.py:369-384)
    assert not np.allclose(res[0]["x0"], res[1]["x0"])
    # the resumed run ends exactly where the uninterrupted one ends, on both ranks
    for r in res:
        assert r["same"] and r["cov_equal"] and r["steps"][0] == r["steps"][1]
        assert r["n_second"] == r["n_full"] > r["n_first"] > 0
        assert r["n_progress"][0] == r["n_progress"][1] >= 1
        assert r["grown"] >= 1  # rows_per_chain: 40 is only the initial size of the store
    # the reference's loader reads both chain files
    from cobaya.output import load_samples

    both = load_samples(str(tmp_path / "b" / "run"), skip=0, combined=False)
    assert len(both) == 2 and [len(c) for c in both] == [r["n_second"] for r in res]


FAILING_RANK = PRELUDE + textwrap.dedent("""
    from cobaya.log import LoggedError

    class Breaks(OracleEngine):
        def advance(self, n):
            if rank == 1 and self._done > 0:
                raise RuntimeError("device fell off the bus")
            super().advance(n)

    plugin.MCMC._engine_factory = Breaks
    try:
        run(copy.deepcopy(info(os.path.join(out, "c", "run"), 5000)), force=True)
        what = "finished"
    except BaseException as e:
        what = type(e).__name__ + ": " + str(e)
    with open(os.path.join(out, f"fail{{rank}}.json"), "w") as f:
        json.dump(dict(rank=rank, what=what), f)
""")


def test_a_failing_rank_ends_every_rank(tmp_path):
    """mpi.py:350-467 / mcmc.py:469: the error of one process surfaces on the others in the
    same iteration of the run loop (packed into the per-launch summary exchange) instead
    of leaving them blocked in the next collective."""
    from tests.refenv import enable_reference

    enable_reference()
    script = tmp_path / "w.py"
    script.write_text(FAILING_RANK.format(root=ROOT, outdir=str(tmp_path)))
    p = _torchrun(script, timeout=300)
    res = {r: json.load(open(tmp_path / f"fail{r}.json")) for r in range(2)}
    assert "device fell off the bus" in res[1]["what"], res
    assert "Another process failed" in res[0]["what"] or "OtherProcess" in res[0]["what"], res


def test_store_comm_implements_what_cobaya_mpi_calls():
    """The mpi4py subset on a c10d store, single process with a HashStore: collectives of
    one rank, and the Isend/iprobe/Recv messages between two communicator objects."""
    import torch.distributed as dist

    from cobaya_b200.distributed import ANY_SOURCE, Status, StoreComm

    store = dist.HashStore()
    a, b = StoreComm(store, 0, 2), StoreComm(store, 1, 2)
    assert a.Get_rank() == 0 and b.Get_size() == 2
    assert not b.iprobe(source=ANY_SOURCE, tag=7)
    a.Isend(np.array([3]), dest=1, tag=7).Test()
    a.Isend(np.array([2]), dest=1, tag=7)
    assert b.iprobe(source=ANY_SOURCE, tag=7) and not b.iprobe(source=ANY_SOURCE, tag=8)
    buf, st = np.empty(1, dtype=int), Status()
    b.Recv(buf, source=ANY_SOURCE, tag=7, status=st)
    assert buf[0] == 3 and st.Get_source() == 0
    b.Recv(buf, source=ANY_SOURCE, tag=7, status=st)
    assert buf[0] == 2 and not b.iprobe(tag=7)
    # collectives need both sides: run rank 1 in a thread
    import threading

    got = {}

    def other():
        got["bc"] = b.bcast(None, root=0)
        got["sc"] = b.scatter(None, root=0)
        b.gather({"r": 1}, root=0)
        got["ag"] = b.allgather("one")
        b.barrier()

    t = threading.Thread(target=other)
    t.start()
    assert a.bcast({"x": 1}, root=0) == {"x": 1}
    assert a.scatter(["zero", "one"], root=0) == "zero"
    assert a.gather({"r": 0}, root=0) == [{"r": 0}, {"r": 1}]
    assert a.allgather("zero") == ["zero", "one"]
    a.barrier()
    t.join(timeout=30)
    assert got == {"bc": {"x": 1}, "sc": "one", "ag": ["zero", "one"]}
