#!/usr/bin/env python
"""
Comparable convergence record (BASELINE.json metric, second half: "R-1 convergence vs ref").

The SAME input -- 64-D correlated Gaussian of SURVEY.md section 8d, uniform priors [-1, 1],
``ref: N(0, 0.001)`` start points drawn through ``Model.get_valid_point``, proposal started
from the diagonal of the target, ``learn_proposal: True`` -- is run with M chains by

* ``--side reference``: the UNMODIFIED reference sampler (cobaya 3.6.2, baseline/_ref), one
  chain per process, M processes under torchrun on the host cores.  mpi4py is not in the
  image; the processes talk through ``cobaya_b200.distributed`` (the mpi4py subset of
  ``cobaya/mpi.py`` on a c10d store), so this is the reference's own multi-chain mode:
  R-1 of means across chains and one learned covariance for all (mcmc.py:773-1032);
* ``--side engine``: the B200 engine through ``cobaya.run.run`` with ``chains_per_gpu: M``
  (the same multi-chain rule on all-reduced sums).

Each side writes R-1 / acceptance / proposals at every convergence check and the final
mean and covariance of the pooled second halves against the analytic truth, in units of
their Monte-Carlo error.  ``--merge`` joins the two files into profiles/.

    python tools/convergence_record.py --side reference --chains 16 --out gpurun_out/conv_ref.json
    python tools/convergence_record.py --side engine --chains 16 --out gpurun_out/conv_eng.json
    python tools/convergence_record.py --merge gpurun_out/conv_ref.json gpurun_out/conv_eng.json \
        --out profiles/r2_convergence_record.json
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p_ in (os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "baseline", "_ref")):
    sys.path.insert(0, p_)

import numpy as np

D = 64


def make_info(sampler_key, extra, max_samples, seed):
    from cobaya_b200.flatmodel import synthetic_gaussian_cov

    cov = synthetic_gaussian_cov(D)
    names = [f"x{i}" for i in range(D)]
    opts = {"covmat": np.diag(np.diag(cov)), "covmat_params": names, "measure_speeds": False,
            "learn_proposal": True, "burn_in": 0, "seed": seed, "max_samples": max_samples,
            "Rminus1_stop": 0.01, "Rminus1_cl_stop": 1e9, "output_every": "1000s"}
    opts.update(extra)
    return {"likelihood": {"gaussian_mixture": {"means": [np.zeros(D)], "covs": [cov],
                                                "input_params": names, "derived": False}},
            "params": {n: {"prior": {"min": -1, "max": 1},
                           "ref": {"dist": "norm", "loc": 0, "scale": 0.001}} for n in names},
            "sampler": {sampler_key: opts}}, cov


def chain_sums(rows):
    """Sufficient statistics of the second half of one chain (what leaves the process)."""
    h = rows[len(rows) // 2:]
    w, X = h[:, 0], h[:, 2:2 + D]
    return {"sw": float(w.sum()), "s1": w @ X, "s2": (X * w[:, None]).T @ X, "rows": len(h)}


def pooled_stats(sums, cov):
    """Mean / covariance of the pooled second halves, and their distance from the truth in
    units of the Monte-Carlo error estimated from the scatter between chains."""
    means = np.array([c["s1"] / c["sw"] for c in sums])
    sw = sum(c["sw"] for c in sums)
    m = sum(c["s1"] for c in sums) / sw
    S = sum(c["s2"] for c in sums) / sw - np.outer(m, m)
    sig = np.sqrt(np.diag(cov))
    mc_err = means.std(axis=0, ddof=1) / np.sqrt(len(sums))  # of the pooled mean
    return {"max_abs_mean_over_sigma": float(np.max(np.abs(m) / sig)),
            "max_abs_mean_over_mc_error": float(np.max(np.abs(m) / mc_err)),
            "rms_mean_over_mc_error": float(np.sqrt(np.mean((m / mc_err) ** 2))),
            "max_rel_err_variance": float(np.max(np.abs(np.diag(S) / np.diag(cov) - 1))),
            "max_abs_err_correlation": float(np.max(np.abs(
                S / np.sqrt(np.outer(np.diag(S), np.diag(S))) -
                cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))))),
            "rows_pooled": int(sum(c["rows"] for c in sums)), "weight_pooled": float(sw)}


REF_WORKER = r"""
import os, sys, json, time
import numpy as np
root, out, chains, max_samples = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sys.path.insert(0, root)
for p in (root + "/oracle/shims", root + "/baseline/_ref"):
    sys.path.insert(0, p)
os.environ["OMP_NUM_THREADS"] = "1"
import logging
import cobaya_b200.distributed as cbd
cbd.init(backend="gloo")
from cobaya import mpi
from cobaya.run import run
sys.path.insert(0, root + "/tools")
import convergence_record as cr
trace = []
def cb(s):
    trace.append((int(s.n()), int(s.n_steps_raw), time.time()))
info, cov = cr.make_info("mcmc", {"callback_function": cb, "callback_every": "10d"}, max_samples, 1)
t0 = time.time()
_, smp = run(info)
wall = time.time() - t0
rows = smp.collection.data.to_numpy()
gathered = mpi.gather((cr.chain_sums(rows), trace, int(smp.n_steps_raw)))
if mpi.is_main_process():
    prog = smp.progress
    res = {"side": "reference", "impl": "cobaya 3.6.2 (unmodified), 1 chain per process, "
           "cobaya.mpi on cobaya_b200.distributed (gloo store)", "chains": chains,
           "wall_s": wall, "converged": bool(smp.converged),
           "proposals_total": int(sum(g[2] for g in gathered)),
           "progress": [{"N": float(r.N), "Rminus1": None if r.Rminus1 is None or not np.isfinite(r.Rminus1) else float(r.Rminus1),
                         "acceptance_rate": float(r.acceptance_rate)} for r in prog.itertuples()],
           "trace_rank0": [list(map(float, t)) for t in gathered[0][1]],
           "final": cr.pooled_stats([g[0] for g in gathered], cov)}
    # proposals (all chains) when each check happened: accepted-steps -> proposals through
    # the per-rank traces (the checks fire when every chain passed a multiple of learn_every)
    json.dump(res, open(out, "w"), indent=1)
"""


def run_reference(args):
    script = os.path.join(ROOT, "gpurun_out", "_conv_ref_worker.py")
    os.makedirs(os.path.dirname(script), exist_ok=True)
    with open(script, "w") as f:
        f.write(REF_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    import socket

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={args.chains}", "--master-addr", "127.0.0.1", "--master-port",
           str(port), script, ROOT, args.out, str(args.chains), str(args.max_samples)]
    with open(args.out + ".log", "w") as lf:
        p = subprocess.run(cmd, env=env, stdout=lf, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(open(args.out + ".log").read()[-6000:])
        raise SystemExit(p.returncode)
    print(open(args.out).read()[:600])


def run_engine(args):
    import logging

    logging.disable(logging.CRITICAL)
    from cobaya.run import run

    import cobaya_b200.plugin  # noqa: F401

    trace = []

    def cb(s):
        trace.append((int(s.n()), int(s.n_steps_raw), time.time()))

    info, cov = make_info("cobaya_b200.plugin.MCMC",
                          {"chains_per_gpu": args.chains, "callback_function": cb,
                           "callback_every": "10d"}, args.max_samples, 1)
    t0 = time.time()
    _, smp = run(info)
    wall = time.time() - t0
    ens = smp._ens
    chains = [chain_sums(ens.chain_rows(c)) for c in range(ens.n_chains_local)]
    res = {"side": "engine", "impl": "cobaya_b200 through cobaya.run.run, chains_per_gpu = "
           f"{args.chains}", "chains": args.chains, "wall_s": wall,
           "converged": bool(smp.converged),
           "proposals_total": int(smp.n_steps_raw) * args.chains,
           "progress": [{"N": float(c.N), "Rminus1": None if c.Rminus1 is None else float(c.Rminus1),
                         "acceptance_rate": float(c.acceptance_rate),
                         "covmat_learned": bool(c.learned)} for c in ens.progress],
           "trace_rank0": [list(map(float, t)) for t in trace],
           "final": pooled_stats(chains, cov)}
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res)[:600])


def merge(args):
    a, b = (json.load(open(f)) for f in args.merge)
    out = {"what": "same input, same start (ref N(0, 0.001)), same multi-chain R-1 rule "
                   "(mcmc.py:773-1032), M chains on both sides; R-1 of means at every "
                   "convergence check against the accepted steps (summed over chains, "
                   "column N of .progress) and the final pooled estimates against the "
                   "analytic truth", a["side"]: a, b["side"]: b}
    json.dump(out, open(args.out, "w"), indent=1)
    for side in (a, b):
        pr = [p for p in side["progress"] if p["Rminus1"] is not None]
        print(side["side"], "checks", len(pr), "last R-1", pr[-1]["Rminus1"] if pr else None,
              "N", pr[-1]["N"] if pr else None, side["final"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", choices=["reference", "engine"])
    ap.add_argument("--merge", nargs=2)
    ap.add_argument("--chains", type=int, default=16)
    ap.add_argument("--max-samples", type=int, default=60000)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    if args.merge:
        merge(args)
    elif args.side == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
