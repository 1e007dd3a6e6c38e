"""
TEST INFRASTRUCTURE ONLY: an object with the interface of ``cobaya_b200.engine.Engine``
whose chains are advanced by the C oracle (``oracle/mcmc_oracle.c``) and whose checkpoint
statistics are numpy restatements.  It lets the CPU test-suite (no GPU here) exercise the
HOST side of the product -- ``EnsembleMCMC``'s run loop, checkpoint rule, growth of the
row store, the plugin's multi-process output and resume -- with world_size 2 on gloo.
The product never imports this module; ``EnsembleMCMC(engine=...)`` is plain dependency
injection and there is no code path that selects it by itself.
"""

from __future__ import annotations

import pickle

import numpy as np

from oracle import oracle as orc


class OracleEngine:
    def __init__(self, fm, n_chains, seed, device=0, chain_id0=0, rows_cap=1024, burn_in=0,
                 rows_want=4096, mem_fraction=0.4):
        self.fm, self.n_chains, self.D = fm, int(n_chains), fm.D
        self.seed, self.chain_id0, self.burn_in = int(seed), int(chain_id0), int(burn_in)
        self.rows_cap = int(rows_want if rows_cap is None else rows_cap)
        self.W = fm.row_width
        self._events = []        # (proposals done, covariance) of every set_covariance
        self._x0 = None
        self._done = 0
        self.grown = 0

    # ---- model / state
    def _new_chains(self, x0):
        self._om = orc.OracleModel(self.fm)
        self._chains = [orc.OracleChain(self._om, self.seed, self.chain_id0 + c, x0[c],
                                        burn_in=self.burn_in) for c in range(self.n_chains)]
        self._rows = [np.zeros((0, self.W)) for _ in range(self.n_chains)]

    def set_state(self, x0):
        self._x0 = np.array(x0, dtype=np.float64).reshape(self.n_chains, self.D)
        for x in self._x0:
            if not np.isfinite(orc.OracleModel(self.fm).logpost(x)[0]):
                raise RuntimeError("initial points have a non-finite log-posterior")
        self._new_chains(self._x0)
        self._done = 0
        self._events = []
        self._cov0 = np.array(self.fm.get_covariance())

    def set_covariance(self, cov):
        self.fm.set_covariance(cov)
        self._om = orc.OracleModel(self.fm)
        for ch in self._chains:
            ch.set_model(self._om)
        self._events.append((self._done, np.array(cov)))

    def advance(self, n):
        for c, ch in enumerate(self._chains):
            room = self.rows_cap - len(self._rows[c])
            rc, rows = ch.advance(int(n), rows_cap=int(n))
            if len(rows) > room:
                raise RuntimeError("rows_cap exceeded (the driver must grow the store first)")
            if rc == 1:
                self._stuck = True
            self._rows[c] = np.concatenate([self._rows[c], rows])
        self._done += int(n)

    def sync(self):
        pass

    def close(self):
        pass

    def get_state(self):
        st = [ch.state() for ch in self._chains]
        return dict(x=np.array([s["x"] for s in st]),
                    logpost=np.array([s["logpost"] for s in st]),
                    weight=np.array([s["weight"] for s in st], np.int64),
                    n_rows=np.array([len(r) for r in self._rows], np.int64),
                    n_accepted=np.array([s["n_accepted"] for s in st], np.int64),
                    flags=np.zeros(self.n_chains, np.uint32))

    def summary(self):
        st = self.get_state()
        return dict(min_rows=int(st["n_rows"].min()), max_rows=int(st["n_rows"].max()),
                    sum_rows=int(st["n_rows"].sum()), n_stuck=int(getattr(self, "_stuck", 0)),
                    n_rows_full=0, n_internal=0, sum_accepted=int(st["n_accepted"].sum()),
                    sum_weight=int(st["weight"].sum()))

    # ---- checkpoint statistics (numpy restatement of cb2_moments / cb2_bounds)
    @property
    def moments_len(self):
        return 3 + self.D + 2 * self.D * self.D

    def moments(self, mode=0, split=4, shift=None, dev_ptr=None, host=True):
        D, DD = self.D, self.D * self.D
        shift = np.zeros(D) if shift is None else np.asarray(shift)
        out = np.zeros(self.moments_len)
        if mode == 0:
            wins = [(r, len(r) // 2, None, len(r)) for r in self._rows]
        else:
            r = self._rows[0]
            cut = len(r) // (1 + split)
            if cut < 2:
                raise RuntimeError("Not enough points in chain to check convergence.")
            wins = [(r, i * cut, (i + 1) * cut - 1, cut) for i in range(1, 1 + split)]
        for rows, a, b, N in wins:
            m, C, acc = orc.chain_window_stats(rows, D, a, b)
            if mode != 0:
                rr = self._rows[0][len(self._rows[0]) // (1 + split):]
                acc = len(rr) / rr[:, 0].sum()
            ms = m - shift
            out[0] += 1; out[1] += N; out[2] += N * acc
            out[3:3 + D] += ms
            out[3 + D:3 + D + DD] += np.outer(ms, ms).ravel()
            out[3 + D + DD:] += (N * C).ravel()
        return out

    def bounds(self, limfrac, mode=0, split=4, shift=None, dev_ptr=None, host=True):
        D = self.D
        shift = np.zeros(D) if shift is None else np.asarray(shift)
        out = np.zeros(1 + 4 * D)
        for rows in self._rows:
            r = rows[len(rows) // 2:]
            out[0] += 1
            for i in range(D):
                v, w = r[:, 2 + i], r[:, 0]
                idx = np.argsort(v, kind="stable")
                cs = np.cumsum(w[idx])
                for k, frac in enumerate((limfrac, 1 - limfrac)):
                    ix = min(np.searchsorted(cs, cs[-1] * frac), len(v) - 1)
                    b = v[idx[ix]] - shift[i]
                    out[1 + 2 * D * k + i] += b
                    out[1 + 2 * D * k + D + i] += b * b
        return out

    # ---- rows
    def rows(self, chain, first=0, n=None):
        r = self._rows[chain][first:]
        return r if n is None else r[:n]

    def rows_bulk(self, first=None, chains=None):
        c0, c1 = (0, self.n_chains) if chains is None else chains
        first = np.zeros(c1 - c0, np.int64) if first is None else np.asarray(first)
        part = [self._rows[c][int(f):] for c, f in zip(range(c0, c1), first)]
        return (np.concatenate(part) if part else np.zeros((0, self.W)),
                np.array([len(p) for p in part], np.int64))

    def mem_info(self):
        return 10**12, 10**12, self.n_chains * self.rows_cap * self.W * 8

    def grow_rows(self, new_cap):
        self.rows_cap = max(self.rows_cap, int(new_cap))
        self.grown += 1

    # ---- resuming: the oracle is deterministic, so the state is "how we got here"
    def export_state(self):
        blob = pickle.dumps(dict(x0=self._x0, done=self._done, events=self._events,
                                 seed=self.seed, id0=self.chain_id0, cov0=self._cov0,
                                 burn_in=self.burn_in))
        return np.frombuffer(blob, dtype=np.uint8).copy()

    def import_state(self, blob, rows=None, counts=None):
        st = pickle.loads(np.asarray(blob, dtype=np.uint8).tobytes())
        if st["seed"] != self.seed or st["id0"] != self.chain_id0:
            raise RuntimeError("snapshot was taken with another seed / chain id range")
        # replay: same proposals, the covariance updates at the same proposal counts
        self.fm.set_covariance(st["cov0"])
        self.burn_in = st["burn_in"]  # a resumed run is configured with burn_in 0 (mcmc.py:1062)
        cap, self.rows_cap = self.rows_cap, 10**9
        self.set_state(st["x0"])
        for at, cov in st["events"]:
            self.advance(at - self._done)
            self.set_covariance(cov)
        self.advance(st["done"] - self._done)
        self.rows_cap = cap
        if rows is not None and counts is not None:  # the restored rows must be these
            got, n = self.rows_bulk()
            assert np.array_equal(n, counts) and np.array_equal(got, np.asarray(rows))
