"""
The D x D part of the checkpoint on the device (cb2_checkpoint_device / cb2_adopt_proposal,
kernels_ckpt.cuh) against the host algebra that restates the reference
(convergence.rminus1_from_sums: mcmc.py:856-889; flatmodel.transforms_from_cov:
proposal.py:226-260, tools.py:761-788):

* R-1, acceptance, mean, W and the candidate transform T to 1e-9 (floating point: device
  Cholesky / Jacobi against LAPACK);
* the device-side repack of the step kernels' constant blocks BIT-EXACT: an engine that
  adopted the candidate on the device and a twin that was handed the same T through
  cb2_set_proposal (host packer) produce identical rows afterwards;
* the driver (EnsembleMCMC) takes the device route by default for D <= 64 and ends at the
  same place as the host route.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _fm(kind):
    from cobaya_b200 import problems
    from cobaya_b200.flatmodel import FlatModel, synthetic_gaussian_cov

    if kind == "c1":
        return problems.config1(64).fm, problems.config1(64).start
    if kind == "d24":
        return problems.config1(24).fm, problems.config1(24).start
    if kind == "d37":      # odd dimension: padded Jacobi, padded fragments
        return problems.config1(37).fm, problems.config1(37).start
    if kind == "blocks":   # two blocks, fast one first in sampler order: i_of_j != identity
        D = 32
        cov = synthetic_gaussian_cov(D)
        fm = FlatModel.gaussian(np.zeros(D), cov, bounds=(-1.0, 1.0),
                                proposal_cov=np.diag(np.diag(cov)),
                                blocks=[list(range(20, 32)), list(range(0, 20))],
                                oversampling=[1, 2])
        return fm, problems.config1(D).start
    if kind == "c3":
        p = problems.config3()
        return p.fm, p.start
    raise KeyError(kind)


def _pair(kind, n_chains=512, warm=600, seed=11):
    from cobaya_b200.engine import Engine

    fm_a, start = _fm(kind)
    fm_b, _ = _fm(kind)
    x0 = start(n_chains, 0)
    engs = []
    for fm in (fm_a, fm_b):
        e = Engine(fm, n_chains=n_chains, seed=seed, chain_id0=0, rows_cap=4096)
        e.set_state(x0)
        e.advance(warm)
        engs.append(e)
    return engs, (fm_a, fm_b)


@pytest.mark.parametrize("kind", ["c1", "d24", "d37", "blocks", "c3"])
def test_device_checkpoint_matches_host_algebra(kind):
    from cobaya_b200.convergence import rminus1_from_sums
    from cobaya_b200.flatmodel import transforms_from_cov

    (a, b), (fm_a, fm_b) = _pair(kind)
    D = fm_a.D
    shift = np.full(D, 0.01)
    sums = a.moments(shift=shift)                  # host copy of what the kernel reads
    host = rminus1_from_sums(sums, D, shift)
    dev = a.checkpoint_device()
    assert dev["success"] and host["success"]
    assert dev["M"] == host["M"] and dev["N"] == host["N"]
    np.testing.assert_allclose(dev["acceptance"], host["acceptance"], rtol=1e-14)
    np.testing.assert_allclose(dev["mean"], host["mean"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(dev["Rminus1"], host["Rminus1"], rtol=RTOL)
    assert 1 <= dev["sweeps"] <= 15
    W = a.checkpoint_cov()
    np.testing.assert_allclose(W, host["W"], rtol=1e-13, atol=1e-18)
    assert dev["proposal_ok"]
    a.adopt_proposal()
    T_dev = a.get_proposal()
    T_host = transforms_from_cov(host["W"], fm_a.i_of_j)
    scale = np.abs(T_host).max()
    np.testing.assert_allclose(T_dev, T_host, rtol=RTOL, atol=RTOL * scale)
    np.testing.assert_array_equal(np.triu(T_dev, 1), 0.0)
    np.testing.assert_allclose(fm_a.get_covariance(), host["W"], rtol=1e-13, atol=1e-18)
    # the twin gets the same transform through the host packer: identical continuation
    b.set_proposal(T_dev)
    for e in (a, b):
        e.advance(300)
    sa, sb = a.get_state(), b.get_state()
    assert not sa["flags"].any()
    np.testing.assert_array_equal(sa["x"], sb["x"])
    np.testing.assert_array_equal(sa["n_rows"], sb["n_rows"])
    np.testing.assert_array_equal(sa["logpost"], sb["logpost"])
    ra, ca = a.rows_bulk()
    rb, cb = b.rows_bulk()
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(ra, rb)
    assert a.last_step_kernel() == b.last_step_kernel()
    a.close(); b.close()


def test_device_checkpoint_reports_a_failed_cholesky():
    """Fewer chains than dimensions: B is singular, d has no zeros but W / d d^T can fail or
    give a huge R-1; with identical chains B = 0 -> d = 0 -> NaN, the reference's
    LinAlgError branch (mcmc.py:872-887)."""
    from cobaya_b200.engine import Engine

    fm, start = _fm("d24")
    x0 = np.repeat(start(1, 0), 16, axis=0)
    e = Engine(fm, n_chains=16, seed=3, chain_id0=0, rows_cap=512)
    e.set_state(x0)
    # no advance: every chain holds the same single row -> zero scatter of the means
    e.advance(1)
    st = e.get_state()
    if st["n_rows"].min() >= 1:
        e.moments(shift=np.zeros(fm.D), host=False)
        dev = e.checkpoint_device()
        from cobaya_b200.convergence import rminus1_from_sums

        host = rminus1_from_sums(e.moments(shift=np.zeros(fm.D)), fm.D, np.zeros(fm.D))
        assert dev["success"] == host["success"]
    e.close()


def test_device_checkpoint_refused_above_64():
    from cobaya_b200.engine import Engine, EngineError
    from cobaya_b200 import problems

    p = problems.config1(72)
    e = Engine(p.fm, n_chains=64, seed=1, chain_id0=0, rows_cap=512)
    e.set_state(p.start(64, 0))
    e.advance(200)
    e.moments(host=False)
    with pytest.raises(EngineError):
        e.checkpoint_device()
    e.close()


def test_driver_takes_the_device_route_and_agrees_with_the_host_route():
    from cobaya_b200 import problems
    from cobaya_b200.mcmc import EnsembleMCMC

    out = {}
    for route in (True, False):
        p = problems.config1(64)
        x0 = p.start(1024, 0)
        ens = EnsembleMCMC(p.fm, x0, {"seed": 5, "max_samples": 400, "learn_proposal": True,
                                      "Rminus1_stop": 0.0, "device_checkpoint": route,
                                      "learn_proposal_Rminus1_max": 1e3,
                                      "learn_proposal_Rminus1_min": 0.0,
                                      "learn_every": "4d"})
        assert ens._use_device_checkpoint() is route
        ens.run()
        out[route] = ens
    dev, host = out[True], out[False]
    assert len(dev.progress) == len(host.progress) >= 1
    # first checkpoint: same rows on both sides -> same numbers up to the algebra's rounding
    np.testing.assert_allclose(dev.progress[0].Rminus1, host.progress[0].Rminus1, rtol=1e-8)
    assert dev.progress[0].N == host.progress[0].N
    assert dev.progress[0].learned == host.progress[0].learned
    assert any(c.learned for c in dev.progress)
    # later checkpoints: transforms differ in the last bits, chains decorrelate slowly;
    # the statistics stay close
    np.testing.assert_allclose(dev.progress[-1].Rminus1, host.progress[-1].Rminus1, rtol=0.2)
    np.testing.assert_allclose(dev.fm.get_covariance(), host.fm.get_covariance(),
                               rtol=0.05, atol=2e-6)
    # default: the device route for D <= 64
    p = problems.config1(64)
    ens = EnsembleMCMC(p.fm, p.start(64, 0), {"seed": 5, "max_samples": 10})
    assert ens._use_device_checkpoint() is True
