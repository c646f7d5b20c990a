"""Planner -> controller reference post-processing.  Golden vectors: tests/golden/planner_refs.npz,
produced by the reference's Map.getGlobalPosition + the SciPy calls of plannerMain.py:257-280 on recorded plans
(tests/golden/make_golden_planrefs.py)."""
import os

import numpy as np
import pytest

import oracle

lp = pytest.importorskip("lpvmpc_b200")
import importlib  # noqa: E402
pp = importlib.import_module("autonomous-racing-lpv-mpp-mpc_b200.postproc")

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "planner_refs.npz"))
TRACK = np.load(os.path.join(os.path.dirname(__file__), "golden", "track.npz"))["L_shape_PointAndTangent"]


def test_filter_constants_are_the_references():
    np.testing.assert_array_equal(pp.ELLIP_B, G["ellip_b"])   # signal.ellip(4, 0.01, 120, 0.125), plannerMain.py:112
    np.testing.assert_array_equal(pp.ELLIP_A, G["ellip_a"])


def test_linear_maps_match_scipy():
    from scipy import signal
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(3)
    for N, dt in ((40, 0.05), (25, 0.05), (63, 1.0 / 30.0)):
        n_out = pp.n_out_for(N, dt)
        W = pp.spline_matrix(N, n_out, N * dt)
        t_in, t_out = np.linspace(0, N * dt, num=N, endpoint=True), np.linspace(0, N * dt, num=n_out, endpoint=True)
        for _ in range(5):
            y = np.cumsum(rng.standard_normal(N)) * 0.3
            want = interp1d(t_in, y, kind="cubic")(t_out)
            np.testing.assert_allclose(W @ y, want, rtol=0, atol=1e-12)
            if n_out > pp.PADLEN:
                F = pp.filtfilt_matrix(pp.ELLIP_B, pp.ELLIP_A, n_out)
                np.testing.assert_allclose(F @ want, signal.filtfilt(pp.ELLIP_B, pp.ELLIP_A, want, padlen=pp.PADLEN), rtol=0, atol=1e-12)


def _raw(x_pred, SS, xyth0):
    N = x_pred.shape[0] - 1
    raw = np.zeros((5, N))
    for i in range(N):
        g = xyth0 if i == 0 else oracle.global_position(TRACK, SS[i], 0.0)
        yaw = g[2] + x_pred[i, 4]
        raw[:, i] = [g[0] - x_pred[i, 3] * np.sin(yaw), g[1] + x_pred[i, 3] * np.cos(yaw), yaw, x_pred[i, 0], x_pred[i, 2] / x_pred[i, 0]]
    return raw


def test_host_pipeline_matches_reference_golden():
    W, Wc = pp.reference_matrices(40, 0.05)
    for c in range(G["refs"].shape[0]):
        raw = _raw(G["x_pred"][c], G["SS"][c], G["xyth0"][c])
        got = np.stack([W @ raw[0], W @ raw[1], W @ raw[2], W @ raw[3], Wc @ raw[4]])
        np.testing.assert_allclose(got, G["refs"][c], rtol=0, atol=1e-11)


@pytest.mark.gpu
def test_device_references_match_reference_golden():
    fleet = lp.PlannerFleet(lp.Map("L_shape"), N=40, max_fleet=64)
    refs, err = fleet.references(G["x_pred"], G["SS"], G["xyth0"])
    assert refs.shape == G["refs"].shape and not err.any()
    np.testing.assert_allclose(refs, G["refs"], rtol=0, atol=1e-10)
    # an arc length outside Curvature()'s domain is flagged, the other plans are untouched
    SS = G["SS"].copy(); SS[3, 7] = -1.0
    refs2, err2 = fleet.references(G["x_pred"], SS, G["xyth0"])
    assert err2[3] == 1 and err2.sum() == 1
    np.testing.assert_allclose(np.delete(refs2, 3, axis=0), np.delete(G["refs"], 3, axis=0), rtol=0, atol=1e-10)
    fleet.close()
