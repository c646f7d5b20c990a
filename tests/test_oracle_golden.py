"""The CPU oracle (oracle/lpv_ref.c) against the golden vectors produced by the reference's own Python
(tests/golden/make_golden.py): track table, curvature lookup, LPV scheduling, QP assembly."""
import os

import numpy as np
import pytest
from scipy import sparse

import oracle

RTOL = 1e-14  # sin/cos and BLAS-vs-loop summation order differ by a few ulp


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _csc(g, p):
    m, n = g[p + "shape"]
    P = sparse.csc_matrix((g[p + "P_data"], g[p + "P_indices"], g[p + "P_indptr"]), shape=(n, n))
    A = sparse.csc_matrix((g[p + "A_data"], g[p + "A_indices"], g[p + "A_indptr"]), shape=(m, n))
    return P, A


def _same_qp(mine, g, p):
    P, A = _csc(g, p)
    assert mine["P"].shape == P.shape and mine["A"].shape == A.shape
    # same sparsity pattern (explicit zeros dropped exactly where scipy drops them)
    assert np.array_equal(mine["P"].indptr, P.indptr) and np.array_equal(mine["P"].indices, P.indices)
    assert np.array_equal(mine["A"].indptr, A.indptr) and np.array_equal(mine["A"].indices, A.indices)
    np.testing.assert_allclose(mine["P"].data, P.data, rtol=RTOL, atol=0)
    np.testing.assert_allclose(mine["A"].data, A.data, rtol=RTOL, atol=1e-18)
    np.testing.assert_allclose(mine["q"], g[p + "q"], rtol=RTOL, atol=1e-15)
    assert np.array_equal(np.isinf(mine["l"]), np.isinf(g[p + "l"]))
    fin = np.isfinite(g[p + "l"])
    np.testing.assert_allclose(mine["l"][fin], g[p + "l"][fin], rtol=RTOL, atol=1e-16)
    np.testing.assert_allclose(mine["u"], g[p + "u"], rtol=RTOL, atol=1e-16)


def test_curvature_matches_reference(golden_dir):
    g = _load(golden_dir, "track.npz")
    track = g["L_shape_PointAndTangent"]
    for s, kap, ok in zip(g["curv_s"], g["curv_kappa"], g["curv_ok"]):
        if ok:
            assert oracle.curvature(s, track) == kap
        else:
            with pytest.raises(TypeError):
                oracle.curvature(s, track)
    with pytest.raises(TypeError):
        oracle.curvature(-0.1, track)
    with pytest.raises(TypeError):
        oracle.curvature(float("nan"), track)


def test_controller_schedule_and_qp(golden_dir):
    g = _load(golden_dir, "controller.npz")
    track = _load(golden_dir, "track.npz")["L_shape_PointAndTangent"]
    assert int(g["n_cases"]) >= 8
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        N, lap, delay = int(g[p + "N"]), int(g[p + "lap"]), int(g[p + "delay"])
        cfg = oracle.make_cfg("controller", N, 1.0 / 30.0, g[p + "Q"], g[p + "R"], g[p + "dR"], track, steering_delay=delay)
        st, A, B, C, err = oracle.ctrl_predict(cfg, g[p + "x"], g[p + "u"], g[p + "vel_ref"], g[p + "curv_ref"], 60.0, lap)
        assert err == 0
        np.testing.assert_allclose(A, g[p + "A"], rtol=RTOL, atol=1e-17)
        np.testing.assert_allclose(B, g[p + "B"], rtol=RTOL, atol=1e-17)
        np.testing.assert_allclose(st, g[p + "states"], rtol=1e-13, atol=1e-15)
        # QP from the reference's own A,B (isolates the build)
        mine = oracle.ctrl_qp(cfg, g[p + "A"], g[p + "B"], g[p + "C"], g[p + "x0"], g[p + "vel_ref"],
                              g[p + "old_steering"], float(g[p + "old_accel"]))
        _same_qp(mine, g, p + "qp_")
        assert "polish=True" in list(g[p + "qp_settings"]) and "verbose=False" in list(g[p + "qp_settings"])
        if delay == 0:
            A2, B2, C2, err = oracle.ctrl_estimate(cfg, g[p + "wu_traj"], g[p + "u"])
            assert err == 0
            np.testing.assert_allclose(A2, g[p + "wu_A"], rtol=RTOL, atol=1e-17)
            np.testing.assert_allclose(B2, g[p + "wu_B"], rtol=RTOL, atol=1e-17)
            mine = oracle.ctrl_qp(cfg, g[p + "wu_A"], g[p + "wu_B"], None, g[p + "x"], g[p + "vel_ref"][:N],
                                  g[p + "old_steering"], float(g[p + "old_accel"]))
            _same_qp(mine, g, p + "wu_qp_")


def test_planner_schedule_and_qp(golden_dir):
    g = _load(golden_dir, "planner.npz")
    track = _load(golden_dir, "track.npz")["L_shape_PointAndTangent"]
    import tests_common as tc
    for i in range(int(g["n_cases"])):
        p = "p%d_" % i
        N = int(g[p + "N"])
        cfg = oracle.make_cfg("planner", N, 1.0 / 20.0, tc.PLAN_Q, tc.PLAN_R, tc.PLAN_DR, track, L_cf=tc.PLAN_L)
        st, A, B, C, err = oracle.plan_predict(cfg, g[p + "x"], g[p + "SS"], g[p + "u"])
        assert err == 0
        np.testing.assert_allclose(A, g[p + "A"], rtol=RTOL, atol=1e-17)
        np.testing.assert_allclose(B, g[p + "B"], rtol=RTOL, atol=1e-17)
        np.testing.assert_allclose(st, g[p + "states"], rtol=1e-13, atol=1e-15)
        mine = oracle.plan_qp(cfg, g[p + "A"], g[p + "B"], None, g[p + "x"], [0.0, 0.0], float(g[p + "max_ey"]))
        _same_qp(mine, g, p + "qp_")
        assert "warm_start=True" in list(g[p + "qp_settings"])
        A2, B2, C2, err = oracle.plan_estimate(cfg, g[p + "wu_traj"], g[p + "wu_uu"])
        assert err == 0
        np.testing.assert_allclose(A2, g[p + "wu_A"], rtol=RTOL, atol=1e-17)
        np.testing.assert_allclose(B2, g[p + "wu_B"], rtol=RTOL, atol=1e-17)
        mine = oracle.plan_qp(cfg, g[p + "wu_A"], g[p + "wu_B"], None, g[p + "x"], [0.0, 0.0], float(g[p + "max_ey"]))
        _same_qp(mine, g, p + "wu_qp_")
