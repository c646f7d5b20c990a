"""the TS-fuzzy (ANFIS) scheduling blend ABC_computation_5SV_new (PathFollowingLPVMPC.py:530-602) and
the polytopic LPV observer GS_LPV_Est (stateEstimator.py:349-492).

tests/golden/aux.npz was produced by the reference's own Python text (tests/golden/make_golden_aux.py; the vertex / gain
tables are absent from the reference repository, so the vectors use seeded random tables of the right shapes).  CPU: the
oracle restatement against the golden vectors.  GPU: the device kernels behind lpvmpc_anfis_abc_* / lpvmpc_observer_step_*
against the golden vectors and the oracle, host and device entry points, ragged sizes and bad arguments."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def test_oracle_anfis_blend_matches_reference_golden(gold):
    A, B, Cc = oracle.anfis_abc(gold["an_sched"], gold["an_A_tab"], gold["an_B_tab"], gold["an_C_tab"], gold["an_bell"])
    np.testing.assert_allclose(A, gold["an_A"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(B, gold["an_B"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(Cc, gold["an_C"], rtol=1e-13, atol=1e-15)


def test_oracle_observer_matches_reference_golden(gold):
    est = gold["ob_est0"]
    for t in range(gold["ob_y"].shape[0]):
        est = oracle.observer_step(est, gold["ob_y"][t], gold["ob_u"][t], gold["ob_lim_ls"], gold["ob_gains_ls"], gold["ob_lim_hs"],
                                   gold["ob_gains_hs"], gold["ob_C"], float(gold["ob_dt"]), int(gold["ob_warm"][t]))
        np.testing.assert_allclose(est, gold["ob_est"][t], rtol=1e-12, atol=1e-13)
        est = gold["ob_est"][t]   # the next tick starts from the reference's own state
    # both polytopes were visited
    vx = gold["ob_est"][:, :, 0]
    assert (vx > gold["ob_lim_ls"][0, 1]).any() and (vx <= gold["ob_lim_ls"][0, 1]).any()


@pytest.mark.gpu
def test_device_anfis_blend(gold):
    lp = pytest.importorskip("lpvmpc_b200")
    import torch
    track = lp.Map("L_shape").PointAndTangent
    s = lp.BatchSolver("controller", 8, lp.workloads.CTRL_DT, track=track, max_batch=64, **lp.workloads.CTRL_TT)
    tabs = (gold["an_A_tab"], gold["an_B_tab"], gold["an_C_tab"], gold["an_bell"])
    r = lp.anfis_abc(s, gold["an_sched"], *tabs)                          # host path
    for k, g in (("A", "an_A"), ("B", "an_B"), ("C", "an_C")):
        np.testing.assert_allclose(r[k], gold[g], rtol=1e-12, atol=1e-14)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    for n in (1, 33, 1000):                                               # device path, ragged sizes, against the oracle
        sched = np.stack([rng.uniform(0.3, 3.5, n), rng.uniform(-0.4, 0.4, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-0.25, 0.25, n),
                          rng.uniform(-1.0, 2.0, n)], axis=1)
        rd = lp.anfis_abc(s, torch.as_tensor(sched).to(dev), *tabs)
        A, B, Cc = oracle.anfis_abc(sched, *tabs)
        np.testing.assert_allclose(rd["A"].cpu().numpy(), A, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(rd["B"].cpu().numpy(), B, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(rd["C"].cpu().numpy(), Cc, rtol=1e-12, atol=1e-14)
    assert lp.anfis_abc(s, np.zeros((0, 5)), *tabs)["A"].shape == (0, 3)
    with pytest.raises(Exception):
        lp.anfis_abc(s, gold["an_sched"], gold["an_A_tab"][:31], *tabs[1:])
    s.close()


@pytest.mark.gpu
def test_device_observer(gold):
    lp = pytest.importorskip("lpvmpc_b200")
    import torch
    track = lp.Map("L_shape").PointAndTangent
    s = lp.BatchSolver("controller", 8, lp.workloads.CTRL_DT, track=track, max_batch=64, **lp.workloads.CTRL_TT)
    tabs = dict(lim_ls=gold["ob_lim_ls"], gains_ls=gold["ob_gains_ls"], lim_hs=gold["ob_lim_hs"], gains_hs=gold["ob_gains_hs"], C_obs=gold["ob_C"])
    dt = float(gold["ob_dt"])
    dev = torch.device("cuda", 0)
    est = gold["ob_est0"]
    est_d = torch.as_tensor(gold["ob_est0"]).to(dev)
    for t in range(gold["ob_y"].shape[0]):
        use = int(gold["ob_warm"][t])
        new = lp.observer_step(s, est, gold["ob_y"][t], gold["ob_u"][t], dt=dt, use_estimate=use, **tabs)            # host path
        np.testing.assert_allclose(new, gold["ob_est"][t], rtol=1e-11, atol=1e-12)
        est = gold["ob_est"][t]
        # device path: the state stays on the device across ticks (own trajectory: compare with the oracle chain)
        est_d = lp.observer_step(s, est_d, torch.as_tensor(gold["ob_y"][t]).to(dev), torch.as_tensor(gold["ob_u"][t]).to(dev), dt=dt,
                                 use_estimate=use, **tabs)
    chain = gold["ob_est0"]
    for t in range(gold["ob_y"].shape[0]):
        chain = oracle.observer_step(chain, gold["ob_y"][t], gold["ob_u"][t], gold["ob_lim_ls"], gold["ob_gains_ls"], gold["ob_lim_hs"],
                                     gold["ob_gains_hs"], gold["ob_C"], dt, int(gold["ob_warm"][t]))
    np.testing.assert_allclose(est_d.cpu().numpy(), chain, rtol=1e-10, atol=1e-11)
    # per-element switch (uint8 / int array) and a bad table shape
    use = np.arange(est.shape[0]) % 2
    new = lp.observer_step(s, gold["ob_est0"], gold["ob_y"][0], gold["ob_u"][0], dt=dt, use_estimate=use, **tabs)
    ref = oracle.observer_step(gold["ob_est0"], gold["ob_y"][0], gold["ob_u"][0], gold["ob_lim_ls"], gold["ob_gains_ls"], gold["ob_lim_hs"],
                               gold["ob_gains_hs"], gold["ob_C"], dt, use)
    np.testing.assert_allclose(new, ref, rtol=1e-11, atol=1e-12)
    with pytest.raises(Exception):
        lp.observer_step(s, gold["ob_est0"], gold["ob_y"][0], gold["ob_u"][0], dt=dt, use_estimate=1, **dict(tabs, gains_ls=gold["ob_gains_ls"][:, :, :8]))
    s.close()
