"""BASELINE configs[3] (Monte-Carlo closed loop) through the C-ABI ``lpvmpc_loop_*``: the device-resident fleet tick
(simulate -> localise -> schedule -> build -> solve) against the CPU oracle's closed loop (oracle/loop_ref.c, itself
pinned to the reference's Python by tests/test_oracle_loop.py).

Tolerances: integer bookkeeping (status, iteration counts summed over all ticks, laps, tick counters) bit-exact;
simulator states / commands within 1e-6 (both sides polish every QP to ~1e-10 and the loop is contractive).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

lp = pytest.importorskip("lpvmpc_b200")
W = lp.workloads


def oracle_cfg(track, N=8):
    return oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_PT["Q"], W.CTRL_PT["R"], W.CTRL_PT["dR"], track)


@pytest.mark.parametrize("swap", [1, 0])
def test_fleet_matches_oracle_loop(swap):
    m = lp.Map("L_shape")
    B, T = 96, 45
    sim0 = lp.fleet_start(B, seed=5, track_map=m)
    sim0[0] = [0.01, 0.0, 0.0, 0.2, 0.0, 0.0, 0.0, 0.0]      # the simulator's own start (vehicleSimulator.py:140-142)
    sim0[1, 0:2] = [30.0, 30.0]                               # off the track: retired at tick 0 with LPVMPC_OFF_TRACK
    sim0[2, 3] = 6.0                                          # vx0 above max_vel with x0 pinned: primal infeasible
    fleet = lp.ClosedLoopFleet(m, N=8, max_fleet=128, swap_ey_epsi=swap)
    fleet.start(sim0)
    # uneven chunks: run(a) + run(b) must be the same fleet as one run(a + b)
    fleet.run(4).run(1).run(T - 5)
    got = fleet.read()
    st = oracle.default_settings(polish=1)
    lc = oracle.loop_cfg(half_width=m.halfWidth, slack=m.slack, swap_ey_epsi=swap)
    ref = oracle.loop_state(sim0, 8)
    oracle.loop_run(oracle_cfg(m.PointAndTangent), st, lc, ref, T, threads=8)
    np.testing.assert_array_equal(got["ctr"], ref["ctr"])
    assert got["ctr"][1, 5] == -22 and got["ctr"][1, 7] == 0
    assert got["ctr"][2, 5] in (-3, 3) and got["ctr"][2, 7] == 0
    alive = got["ctr"][:, 5] == 0
    assert alive.sum() == B - 2
    np.testing.assert_array_equal(got["stat"][:, [0, 1, 3]], ref["stat"][:, [0, 1, 3]])
    np.testing.assert_allclose(got["stat"][:, 2], ref["stat"][:, 2], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got["sim"], ref["sim"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["cmd"], ref["cmd"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["local"][alive], ref["local"][alive], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["u_pred"][alive], ref["u_pred"][alive], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["x_pred"][alive], ref["x_pred"][alive], rtol=0, atol=1e-6)
    fleet.close()


def test_fleet_device_start_and_restart_are_deterministic():
    import torch
    m = lp.Map("L_shape")
    sim0 = lp.fleet_start(512, seed=9, track_map=m)
    fleet = lp.ClosedLoopFleet(m, N=8, max_fleet=512)
    fleet.start(sim0).run(30)
    a = fleet.read()
    t = torch.as_tensor(sim0).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    fleet.start(t).run(30, stream=stream)
    b = fleet.read()          # waits for the caller's stream through the handle's event
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    fleet.close()


def test_full_fleet_one_lap_properties():
    """8,192 vehicles (one GPU's share of configs[3]) x 600 ticks: nobody is retired, every tick is SOLVED, every car
    stays within the track and completes its lap; the per-tick API path (host loop over BatchSolver.solve with the
    oracle's simulator) is covered by test_gpu_closed_loop.py."""
    m = lp.Map("L_shape")
    B, T = 8192, 600
    fleet = lp.ClosedLoopFleet(m, N=8, max_fleet=B)
    fleet.start(lp.fleet_start(B, seed=2, track_map=m)).run(T)
    r = fleet.read(("ctr", "stat", "sim", "local"))
    ctr, stat = r["ctr"], r["stat"]
    assert (ctr[:, 5] == 0).all() and (ctr[:, 7] == T).all()
    assert (stat[:, 0] == T).all()
    assert stat[:, 2].max() < 0.12
    assert (ctr[:, 1] >= 1).all() and (stat[:, 3] >= 0).all()
    assert np.isfinite(r["sim"]).all()
    assert 0.9 < np.median(r["local"][:, 0]) < 1.1            # settled at the 1 m/s reference
    fleet.close()


def test_planner_fleet_matches_oracle_loop():
    """SURVEY 8f rows 2-3 (planner main loop, plannerMain.py:128-224) through lpvmpc_plan_loop_*: the Testing-mode start
    [1, 0, 0, 0, 0] at s = 0 plus perturbed starts spread over the track, 12 ticks, against the oracle's planner loop
    (pinned to the reference's LPV_MPC_Planner by tests/test_oracle_loop.py)."""
    m = lp.Map("L_shape")
    rng = np.random.default_rng(21)
    B, T, N = 40, 12, 40
    x0 = np.stack([rng.uniform(1.0, 2.0, B), rng.normal(0, 0.01, B), rng.normal(0, 0.05, B), rng.normal(0, 0.02, B), rng.normal(0, 0.02, B)], axis=1)
    s0 = rng.uniform(0.0, 18.0, B)
    x0[0] = [1.0, 0.0, 0.0, 0.0, 0.0]; s0[0] = 0.0
    x0[1, 0] = 0.5                          # below min_vel with x0 pinned: primal infeasible, retired at tick 0
    fleet = lp.PlannerFleet(m, N=N, max_fleet=64, max_ey=0.2)
    cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], m.PointAndTangent, L_cf=W.PLAN["L_cf"])
    ref = oracle.plan_loop_state(x0, N, s0)
    fleet.start(x0, s0)
    # Statuses, iteration counts and tick counters must be identical throughout.  Plans agree to 1e-6 over the first
    # ticks; they are only solved to OSQP's eps = 1e-3 (polish rarely succeeds on planner QPs) and every tick re-plans
    # from the previous plan, so round-off level differences grow: within 1e-3 (the solver's own accuracy) after 12.
    for ticks, tol in ((4, 2e-6), (T - 4, 1e-3)):
        got = fleet.run(ticks).read()
        oracle.plan_loop_run(cfg, oracle.default_settings(polish=1), ref, ticks, max_ey=0.2, threads=8)
        np.testing.assert_array_equal(got["ctr"], ref["ctr"])
        np.testing.assert_array_equal(got["stat"], ref["stat"])
        alive = got["ctr"][:, 3] == 0
        np.testing.assert_allclose(got["SS"][alive], ref["SS"][alive], rtol=0, atol=tol)
        np.testing.assert_allclose(got["x_pred"][alive], ref["x_pred"][alive], rtol=0, atol=tol)
        np.testing.assert_allclose(got["u_pred"][alive], ref["u_pred"][alive], rtol=0, atol=tol)
    assert got["ctr"][1, 3] in (-3, 3) and got["ctr"][1, 0] == 0
    assert alive.sum() >= B - 4 and (got["ctr"][alive, 0] == T).all()
    fleet.close()
