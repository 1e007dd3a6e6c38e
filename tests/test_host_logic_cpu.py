"""CPU tests of the host-side driver logic (no GPU): option units, learning gate, the
sum formulation of the checkpoint, and the N>1 exchange with the gloo backend
(world_size 2)."""

import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tests.util import ROOT


def _free_port():
    import socket

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def test_number_with_units():
    from cobaya_b200.mcmc import NumberWithUnits

    n = NumberWithUnits("40d", "d", scale=7)
    assert n.value == 280 and n.unit == "d" and n.unit_value == 40
    assert NumberWithUnits(15, "d", scale=7).value == 15
    assert NumberWithUnits("d", "d", scale=3).value == 3
    assert NumberWithUnits("60s", "s").unit == "s"
    assert NumberWithUnits(np.inf, "d", dtype=float, scale=3).value == np.inf


def test_learning_gate_follows_reference():
    """mcmc.py:1009-1030"""
    from cobaya_b200.mcmc import MCMC_DEFAULTS, decide_learning

    o = dict(MCMC_DEFAULTS)
    assert decide_learning(o, 1.0, False)[0]
    assert not decide_learning(o, 3.0, False)[0]          # > Rminus1_max (2)
    assert not decide_learning(o, 1.0, True)[0]           # converged
    o["learn_proposal_Rminus1_min"] = 0.5
    assert not decide_learning(o, 0.1, False)[0]
    o["learn_proposal"] = False
    assert not decide_learning(o, 1.0, False)[0]


def test_mcmc_defaults_cover_reference_yaml_keys():
    """Appendix B of SURVEY.md: every mcmc.yaml key is known to the standalone driver."""
    from cobaya_b200.mcmc import MCMC_DEFAULTS

    keys = ["burn_in", "max_tries", "covmat", "covmat_params", "proposal_scale",
            "output_every", "learn_every", "temperature", "learn_proposal",
            "learn_proposal_Rminus1_max", "learn_proposal_Rminus1_max_early",
            "learn_proposal_Rminus1_min", "max_samples", "Rminus1_stop", "Rminus1_cl_stop",
            "Rminus1_cl_level", "Rminus1_single_split", "measure_speeds", "oversample_power",
            "oversample_thin", "drag", "blocking", "callback_function", "callback_every",
            "seed", "check_every", "oversample", "drag_limits"]
    assert set(keys) <= set(MCMC_DEFAULTS)


def local_sums(rows_per_chain, D, shift):
    """numpy restatement of cb2_moments (mode HALVES) for a list of chains' rows."""
    from oracle import oracle as orc

    DD = D * D
    out = np.zeros(3 + D + 2 * DD)
    for rows in rows_per_chain:
        n = len(rows)
        m, C, a = orc.chain_window_stats(rows, D, n // 2)
        ms = m - shift
        out[0] += 1; out[1] += n; out[2] += n * a
        out[3:3 + D] += ms
        out[3 + D:3 + D + DD] += np.outer(ms, ms).ravel()
        out[3 + D + DD:] += (n * C).ravel()
    return out


def _make_chains(n_chains, n, seed=3):
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov
    from oracle import oracle as orc

    D = 5
    cov = synthetic_gaussian_cov(D)
    fm = FlatModel.gaussian(np.full(D, 0.2), cov, proposal_cov=cov)
    om = orc.OracleModel(fm)
    rows = []
    for c in range(n_chains):
        ch = orc.OracleChain(om, seed, c, np.full(D, 0.2))
        rows.append(ch.advance(n)[1])
    return D, rows


def test_sum_formulation_equals_reference_gather_formulation():
    from cobaya_b200.convergence import rminus1_from_sums
    from oracle import oracle as orc

    D, rows = _make_chains(6, 500)
    shift = np.full(D, 0.19)
    res = rminus1_from_sums(local_sums(rows, D, shift), D, shift)
    st = [orc.chain_window_stats(r, D, len(r) // 2) for r in rows]
    R, W = orc.rminus1_from_chain_stats([len(r) for r in rows], [s[0] for s in st],
                                        [s[1] for s in st])
    np.testing.assert_allclose(res["Rminus1"], R, rtol=1e-8)
    np.testing.assert_allclose(res["W"], W, rtol=1e-12)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    from tests.test_host_logic_cpu import _make_chains, local_sums
    from cobaya_b200.mcmc import TorchDist
    from cobaya_b200.convergence import rminus1_from_sums
    dist.init_process_group("gloo")
    td = TorchDist()
    D, rows = _make_chains(6, 500)
    mine = rows[td.rank * 3:(td.rank + 1) * 3]          # chains sharded over ranks
    shift = np.full(D, 0.19)
    tot = td.all_reduce_sum(local_sums(mine, D, shift))
    res = rminus1_from_sums(tot, D, shift)
    # the per-launch summary: one all-gather of the ranks' int64 vectors
    allv = td.all_gather_i64([min(len(r) for r in mine), max(len(r) for r in mine),
                              sum(len(r) for r in mine)])
    mn, mx, sm = [allv[:, 0].min()], [allv[:, 1].max()], [allv[:, 2].sum()]
    # samples(combined=True) / to_getdist: per-rank row arrays gathered on every rank
    gathered = td.all_gather_object([np.asarray(r) for r in mine])
    n_gathered = [sum(len(r) for r in part) for part in gathered]
    # one file per rank: the two ranks' stdout lines can interleave
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                           f"rank{{td.rank}}.json"), "w") as f:
        json.dump(dict(rank=td.rank, R=res["Rminus1"], N=res["N"],
                       W00=float(res["W"][0, 0]), mn=int(mn[0]),
                       mx=int(mx[0]), sm=int(sm[0]), n_gathered=n_gathered), f)
    dist.destroy_process_group()
""")


def test_world_size_2_gloo_allreduce_gives_identical_verdict_on_every_rank(tmp_path):
    from cobaya_b200.convergence import rminus1_from_sums

    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    import json

    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    assert res[0]["R"] == res[1]["R"] and res[0]["W00"] == res[1]["W00"]   # bit-identical
    D, rows = _make_chains(6, 500)
    shift = np.full(D, 0.19)
    single = rminus1_from_sums(local_sums(rows, D, shift), D, shift)
    np.testing.assert_allclose(res[0]["R"], single["Rminus1"], rtol=1e-10)
    assert res[0]["N"] == single["N"] == sum(len(r) for r in rows)
    assert res[0]["mn"] == min(len(r) for r in rows) and res[0]["mx"] == max(len(r) for r in rows)
    # every rank holds every rank's chains after the gather, in rank order
    expect = [sum(len(r) for r in rows[:3]), sum(len(r) for r in rows[3:])]
    assert res[0]["n_gathered"] == expect and res[1]["n_gathered"] == expect


def test_cabi_exports_every_declared_symbol():
    """The in-tree CUDA library loads on a CPU-only box and exports every function
    include/cobaya_b200.h declares (no compute calls without a GPU)."""
    import re

    from cobaya_b200 import _cabi

    hdr = open(os.path.join(ROOT, "include", "cobaya_b200.h")).read()
    declared = set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", hdr))
    L = _cabi.load()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert declared <= set(_cabi.EXPORTS) | {"cb2_engine"}
    assert L.cb2_abi_version() >= 1


def test_engine_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cobaya_b200.engine import Engine, EngineError
    from cobaya_b200.flatmodel import FlatModel

    fm = FlatModel.gaussian(np.zeros(2), np.eye(2) * 0.01, proposal_cov=np.eye(2) * 0.01)
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(fm, n_chains=2, seed=1)
