"""The oracle's closed-loop pieces (oracle/loop_ref.c) against vectors produced by the REFERENCE'S OWN Python
(tests/golden/closed_loop.npz, generator tests/golden/make_golden_loop.py): Simulator.f, Map.getLocalPosition,
Map.getGlobalPosition, predicted_vectors_generation and 60 ticks of the controller main loop."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "closed_loop.npz"))
T = np.load(os.path.join(os.path.dirname(__file__), "golden", "track.npz"))
TRACK = T["L_shape_PointAndTangent"]


def test_simulator_step_matches_reference():
    st, u, nxt = G["sim_state"], G["sim_u"], G["sim_next"]
    for i in range(st.shape[0]):
        got = oracle.sim_f(st[i], u[i])
        np.testing.assert_allclose(got, nxt[i], rtol=1e-14, atol=1e-15)


def test_local_position_matches_reference():
    pts, ref = G["lp_in"], G["lp_out"]
    hw, slack = float(G["halfWidth"]), float(G["slack"])
    n_off = 0
    for p, r in zip(pts, ref):
        s, ey, epsi, flag = oracle.local_position(TRACK, hw, slack, *p)
        assert flag == int(r[3])
        n_off += flag == 0
        np.testing.assert_allclose([s, ey, epsi], r[:3], rtol=0, atol=2e-12)
    assert n_off > 5  # the off-track sentinel path is exercised


def test_global_position_matches_reference_and_round_trips():
    hw, slack = float(G["halfWidth"]), float(G["slack"])
    for s, ey, r in zip(G["gp_s"], G["gp_ey"], G["gp_out"]):
        x, y, th = oracle.global_position(TRACK, s, ey)
        np.testing.assert_allclose([x, y, th], r, rtol=0, atol=1e-13)
        s2, ey2, epsi2, flag = oracle.local_position(TRACK, hw, slack, x, y, th)
        if flag and 1e-9 < s < float(T["L_shape_TrackLength"]) - 1e-9:
            # unityTestChangeOfCoordinates (trackInitialization.py:433-460): 1e-8 round trip
            assert abs(s2 - s) < 1e-8 and abs(ey2 - ey) < 1e-8 and abs(epsi2) < 1e-8


def test_guess_vectors_match_reference():
    N = G["guess_xx"].shape[0]
    xx, uu = np.zeros((20, 6)), np.zeros((20, 2))
    oracle.lib().loop_ref_guess(oracle._dp(np.ascontiguousarray(G["guess_local"])), 20, oracle._dp(xx), oracle._dp(uu))
    np.testing.assert_array_equal(xx[:N], G["guess_xx"][:N])
    np.testing.assert_array_equal(uu[:N], G["guess_uu"][:N])


@pytest.mark.parametrize("swap", [1, 0])
def test_closed_loop_matches_reference_loop(swap):
    """60 ticks of reference controller class + Map + Simulator (QP solved by the oracle's OSQP at the stub seam)
    against the oracle's own C loop: states, local states, commands, statuses, iteration counts."""
    p = "loop%d_" % swap
    sim, local, cmd, status = G[p + "sim"], G[p + "local"], G[p + "cmd"], G[p + "status"]
    N = G[p + "upred"].shape[1]
    Q = np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0])
    cfg = oracle.make_cfg("controller", N, 1.0 / 30.0, Q, 0.25 * np.eye(2), 37.5 * np.array([1.3, 1.0]), TRACK)
    st = oracle.default_settings(polish=1)
    lc = oracle.loop_cfg(half_width=float(G["halfWidth"]), slack=float(G["slack"]), swap_ey_epsi=swap)
    state = oracle.loop_state(sim[:1], N)
    for t in range(sim.shape[0]):
        np.testing.assert_allclose(state["sim"][0], sim[t], rtol=0, atol=1e-9)
        oracle.loop_run(cfg, st, lc, state, 1)
        np.testing.assert_allclose(state["local"][0], local[t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(state["cmd"][0], cmd[t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(state["u_pred"][0], G[p + "upred"][t], rtol=0, atol=1e-9)
        assert state["ctr"][0, 3] == status[t, 0] and state["ctr"][0, 4] == status[t, 1]
    np.testing.assert_allclose(state["sim"][0], G[p + "final_sim"], rtol=0, atol=1e-9)
    assert state["ctr"][0, 7] == sim.shape[0] and state["ctr"][0, 5] == 0


def test_full_lap_stays_on_track():
    """One lap of the L_shape track at 1 m/s (about 552 ticks): every tick SOLVED, |ey| small."""
    Q = np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0])
    cfg = oracle.make_cfg("controller", 8, 1.0 / 30.0, Q, 0.25 * np.eye(2), 37.5 * np.array([1.3, 1.0]), TRACK)
    st = oracle.default_settings(polish=1)
    for swap in (1, 0):
        lc = oracle.loop_cfg(half_width=float(G["halfWidth"]), slack=float(G["slack"]), swap_ey_epsi=swap)
        state = oracle.loop_state(np.array([[0.01, 0.0, 0.0, 0.2, 0.0, 0.0, 0.0, 0.0]]), 8)
        oracle.loop_run(cfg, st, lc, state, 640)
        ctr, stat = state["ctr"][0], state["stat"][0]
        assert ctr[5] == 0 and ctr[7] == 640, (swap, ctr)
        assert ctr[1] == 1 and 500 < stat[3] < 620, (swap, ctr, stat)   # one lap completed
        assert stat[2] < 0.15, (swap, stat)


# ------------------------------------------------------------------------------------------------ planner loop
GP = np.load(os.path.join(os.path.dirname(__file__), "golden", "planner_loop.npz"))
PLAN_Q = -np.diag([-0.000000000000088, -9.703658572659423, -0.5, 0.000000000213635, -0.153591566469547])
PLAN_L = -np.array([1.00702414775175, 0.187661946033823, -0.0, 0.0, -0.0329493219494661])


def test_planner_guess_matches_reference():
    xx = oracle.plan_guess(GP["guess_x0"], 40, 0.2, 0.05)
    np.testing.assert_allclose(xx, GP["guess_xx"], rtol=1e-15, atol=1e-16)


@pytest.mark.parametrize("case", [0, 1])
def test_planner_loop_matches_reference_loop(case):
    """25 ticks of the reference's LPV_MPC_Planner driven by the planner main loop (QP solved by the oracle's OSQP at the
    stub seam) against the oracle's own C loop: plans, inputs, arc lengths, statuses, iteration counts."""
    p = "loop%d_" % case
    xp, up, SS, status = GP[p + "xpred"], GP[p + "upred"], GP[p + "SS"], GP[p + "status"]
    N = up.shape[1]
    cfg = oracle.make_cfg("planner", N, 1.0 / 20.0, PLAN_Q, np.diag([0.8, 0.0]), np.array([6.0, 6.0]), TRACK, L_cf=PLAN_L)
    st = oracle.default_settings(polish=1)
    state = oracle.plan_loop_state(GP[p + "x0"][None, :], N)
    for t in range(xp.shape[0]):
        oracle.plan_loop_run(cfg, st, state, 1, max_ey=0.2)
        assert state["ctr"][0, 1] == status[t, 0] and state["ctr"][0, 2] == status[t, 1], (t, state["ctr"][0], status[t])
        np.testing.assert_allclose(state["x_pred"][0], xp[t], rtol=0, atol=1e-8)
        np.testing.assert_allclose(state["u_pred"][0], up[t], rtol=0, atol=1e-8)
        np.testing.assert_allclose(state["SS"][0], SS[t], rtol=0, atol=1e-8)
    assert state["ctr"][0, 0] == xp.shape[0] and state["ctr"][0, 3] == 0
