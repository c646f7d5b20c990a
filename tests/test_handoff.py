"""Controller <- planner hand-off (Body_Frame_Errors, reference windowing).

Golden vectors: tests/golden/handoff.npz — the reference's own lines controllerMain.py:200-243 exec'd as text together
with its Body_Frame_Errors and wrap (tests/golden/make_golden_handoff.py), for max_window = 0 (as in the file) and 3.

CPU tests: the host-side index state machine (ReferenceWindow) bit-exact, the oracle restatement (oracle/loop_ref.c:
track_inputs_ref) within 1e-14 (libm vs numpy sin/cos).  GPU tests: the kernel through the C-ABI against the golden
vectors (1e-12) and the oracle, and the full chain hand-off -> LPVPrediction(lap=1) -> solve against the oracle solve.
"""
import importlib
import os

import numpy as np
import pytest

import oracle

lp = pytest.importorskip("lpvmpc_b200")
ho = importlib.import_module("autonomous-racing-lpv-mpp-mpc_b200.handoff")
W = lp.workloads

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "handoff.npz"))
N, DT = 8, 1.0 / 30.0
WIN = ("x_ref", "y_ref", "yaw_ref", "vel_ref", "curv_ref")


def _windows(tag):
    """[T,5,N] latched windows of the golden run (what the reference's variables held at each tick)."""
    return np.stack([G["%s_%s" % (tag, k)] for k in WIN], axis=1)


@pytest.mark.parametrize("tag,mw", [("w0", 0), ("w3", 3)])
def test_reference_window_state_machine(tag, mw):
    w = ho.ReferenceWindow(N, max_window=mw)
    want = _windows(tag)
    for t in range(want.shape[0]):
        assert w.index == G[tag + "_index_before"][t]
        got = w.update(*G[tag + "_msg"][t])
        for i in range(5):
            np.testing.assert_array_equal(got[i], want[t, i])
    if mw == 0:   # the file's value: a fresh window every second tick
        np.testing.assert_array_equal(G["w0_fresh"], np.arange(40) % 2 == 0)


@pytest.mark.parametrize("tag", ["w0", "w3"])
def test_oracle_matches_reference_body_frame_errors(tag):
    r = oracle.track_inputs(G[tag + "_gstate"], G[tag + "_s_prev"], _windows(tag), N, DT, lap=G[tag + "_lap"])
    np.testing.assert_allclose(r["x0"], G[tag + "_local"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(r["ex"], G[tag + "_ex"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(r["x0"][:, 4], G[tag + "_s_after"], rtol=0, atol=1e-14)
    np.testing.assert_array_equal(r["vel_ref"][:, :N], G[tag + "_vel_ref"])
    np.testing.assert_array_equal(r["vel_ref"][:, N], G[tag + "_vel_ref"][:, -1])
    np.testing.assert_array_equal(r["curv_ref"], G[tag + "_curv_ref"])
    assert np.abs(r["x0"][:, 3]).max() <= np.pi    # yaw carried 1..3 laps (+ one extra turn on some ticks): wrapped


def test_oracle_window_offsets():
    """index > 0 on the full message == index 0 on the pre-cut window."""
    m = G["w3_msg"][:12]
    idx = np.arange(12, dtype=np.int32) % 4
    a = oracle.track_inputs(G["w3_gstate"][:12], G["w3_s_prev"][:12], m, N, DT, lap=G["w3_lap"][:12], index=idx)
    cut = np.stack([m[b, :, idx[b]:idx[b] + N] for b in range(12)])
    b_ = oracle.track_inputs(G["w3_gstate"][:12], G["w3_s_prev"][:12], cut, N, DT, lap=G["w3_lap"][:12])
    for k in a:
        np.testing.assert_array_equal(a[k], b_[k])


# ---------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def solver():
    s = lp.BatchSolver("controller", N, DT, track=lp.Map("L_shape").PointAndTangent, max_batch=512, **W.CTRL_TT)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["w0", "w3"])
def test_gpu_handoff_matches_reference(solver, tag):
    r = ho.track_inputs(solver, G[tag + "_gstate"], G[tag + "_s_prev"], _windows(tag), lap=G[tag + "_lap"])
    np.testing.assert_allclose(r["x0"], G[tag + "_local"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r["ex"], G[tag + "_ex"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(r["vel_ref"][:, :N], G[tag + "_vel_ref"])
    np.testing.assert_array_equal(r["vel_ref"][:, N], G[tag + "_vel_ref"][:, -1])
    np.testing.assert_array_equal(r["curv_ref"], G[tag + "_curv_ref"])


def _random_case(B, seed):
    rng = np.random.default_rng(seed)
    msgs = np.concatenate([G["w0_msg"], G["w3_msg"]])
    m = msgs[rng.integers(0, msgs.shape[0], B)]
    idx = rng.integers(0, 6, B).astype(np.int32)
    lap = rng.integers(0, 4, B).astype(np.int32)
    head = m[np.arange(B), :, idx]                                    # [B,5] reference pose at the window start
    g = np.stack([rng.uniform(0.6, 2.5, B), rng.normal(0, 0.1, B), rng.normal(0, 0.5, B), head[:, 0] + rng.normal(0, 0.05, B),
                  head[:, 1] + rng.normal(0, 0.05, B), head[:, 2] + rng.normal(0, 0.1, B) + 2 * np.pi * lap], axis=1)
    return g, rng.uniform(0, 19, B), m, lap, idx


@pytest.mark.gpu
def test_gpu_handoff_matches_oracle_host_and_device_paths(solver):
    import torch
    g, sp, m, lap, idx = _random_case(300, 11)
    want = oracle.track_inputs(g, sp, m, N, DT, lap=lap, index=idx)
    got = ho.track_inputs(solver, g, sp, m, lap=lap, index=idx)
    dev = torch.device("cuda", 0)
    got_d = ho.track_inputs(solver, torch.from_numpy(g).to(dev), torch.from_numpy(sp).to(dev), torch.from_numpy(m).to(dev),
                            lap=torch.from_numpy(lap).to(dev), index=torch.from_numpy(idx).to(dev))
    torch.cuda.synchronize()
    for k in ("x0", "vel_ref", "curv_ref", "ex"):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(got_d[k].cpu().numpy(), got[k])
    # optional pointers: no lap (= 0), no index (= 0), no ex
    got0 = ho.track_inputs(solver, g, sp, m)
    want0 = oracle.track_inputs(g, sp, m, N, DT)
    np.testing.assert_allclose(got0["x0"], want0["x0"], rtol=0, atol=1e-12)
    assert ho.track_inputs(solver, g[:0], sp[:0], m[:0])["x0"].shape == (0, 6)


@pytest.mark.gpu
def test_gpu_handoff_argument_errors(solver):
    g, sp, m, lap, idx = _random_case(4, 3)
    with pytest.raises(lp.NativeError):
        ho.track_inputs(solver, g, sp, m[:, :, :N + 2], index=np.full(4, 3, np.int32))     # window runs off the message
    with pytest.raises(lp.NativeError):
        ho.track_inputs(solver, g, sp, m, index=np.array([0, -1, 0, 0], np.int32))
    p = lp.BatchSolver("planner", 40, 0.05, track=lp.Map("L_shape").PointAndTangent, max_batch=8, **W.PLAN)
    with pytest.raises(lp.NativeError):
        ho.track_inputs(p, g, sp, m)
    p.close()


@pytest.mark.gpu
def test_gpu_tracking_tick_matches_oracle(solver):
    """controllerMain.py:196-243 + 361-363: hand-off -> LPVPrediction(LocalState, uPred, vel_ref, curv_ref, Cf, Lap >= 1)
    -> solve, device chain against the oracle chain: status / iterations identical, commands within 1e-6."""
    B = 128
    g, sp, m, lap, idx = _random_case(B, 21)
    lap = np.maximum(lap, 1).astype(np.int32)
    g[:, 5] += 2 * np.pi * (lap - np.round((g[:, 5] - m[np.arange(B), 2, idx]) / (2 * np.pi)))
    rng = np.random.default_rng(22)
    u_prev = np.stack([np.clip(rng.uniform(-0.1, 0.1, (B, 1)) + 0.01 * np.cumsum(rng.standard_normal((B, N)), axis=1), -0.249, 0.249),
                       np.repeat(rng.uniform(-0.5, 1.5, (B, 1)), N, axis=1)], axis=2)
    h = ho.track_inputs(solver, g, sp, m, lap=lap, index=idx)
    r = solver.solve(h["x0"], u_prev=u_prev, vel_ref=h["vel_ref"], curv_ref=h["curv_ref"], lap=lap, u_old=u_prev[:, 0].copy())
    ref_in = oracle.track_inputs(g, sp, m, N, DT, lap=lap, index=idx)
    cfg = oracle.make_cfg("controller", N, DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], lp.Map("L_shape").PointAndTangent)
    o = oracle.ctrl_batch(cfg, oracle.default_settings(polish=1), ref_in["x0"], u_prev, ref_in["vel_ref"], ref_in["curv_ref"], lap,
                          u_prev[:, 0].copy(), threads=8)
    np.testing.assert_array_equal(r.status, o["status"])
    np.testing.assert_array_equal(r.iters, o["iters"])
    ok = o["status"] == 1
    assert ok.mean() > 0.9
    np.testing.assert_allclose(r.u_pred[ok], o["uPred"][ok], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r.x_pred[ok], o["xPred"][ok], rtol=0, atol=1e-6)
