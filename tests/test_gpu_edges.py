"""Edge cases of the batched path through the C-ABI: empty and single-problem batches, ragged batch sizes around the
4-QPs-per-warp grouping, the shortest and longest horizons a handle accepts, over-sized batches, the N = 20 variant of
BASELINE configs[0], failing Curvature look-ups inside a batch, warm-started statuses that must not leak between
problems sharing a warp."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

lp = pytest.importorskip("lpvmpc_b200")
W = lp.workloads
KEYS = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")


@pytest.fixture(scope="module")
def track():
    return lp.Map("L_shape").PointAndTangent


def _check_ctrl(track, N, w, r, idx, tune=None):
    tune = tune or W.CTRL_TT
    cfg = oracle.make_cfg("controller", N, W.CTRL_DT, tune["Q"], tune["R"], tune["dR"], track)
    st = oracle.default_settings(polish=1)
    for b in idx:
        o = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                              curv_ref=w["curv_ref"][b], lap=int(w["lap"][b]), old_steering=[w["u_old"][b, 0]],
                              old_accel=float(w["u_old"][b, 1]))
        want = -20 if o["sched_err"] else o["status"]
        assert int(r.status[b]) == want, (b, r.status[b], want)
        if want in (1, 2, -2):
            assert int(r.iters[b]) == o["iter"], (b, r.iters[b], o["iter"])
            np.testing.assert_allclose(r.u_pred[b], o["uPred"], rtol=0, atol=1e-4)
            np.testing.assert_allclose(r.x_pred[b], o["xPred"], rtol=0, atol=1e-4)
        else:
            assert np.isnan(r.u_pred[b]).all()


@pytest.mark.parametrize("B", [0, 1, 2, 3, 5, 63, 65])
def test_ragged_batch_sizes(track, B):
    N = 8
    w = W.controller_batch(max(B, 1), N, seed=100 + B)
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=80, **W.CTRL_TT)
    r = s.solve(w["x0"][:B], **{k: w[k][:B] for k in KEYS})
    assert r.status.shape == (B,) and r.u_pred.shape == (B, N, 2) and r.x_pred.shape == (B, N + 1, 6)
    assert s.info()["kernel_launches"] == (0 if B == 0 else (1 if B < 8 else 2))   # batches of 8+ controller QPs: order kernel + solve
    _check_ctrl(track, N, w, r, range(B))
    s.close()


def test_batch_larger_than_max_batch_is_refused(track):
    w = W.controller_batch(9, 8, seed=1)
    s = lp.BatchSolver("controller", 8, W.CTRL_DT, track=track, max_batch=8, **W.CTRL_TT)
    with pytest.raises(lp.NativeError):
        s.solve(w["x0"], **{k: w[k] for k in KEYS})
    s.close()


@pytest.mark.parametrize("N", [2, 3, 9, 20])
def test_horizon_range_controller(track, N):
    """N = 2 is the shortest horizon lpvmpc_create accepts; 9 is the first horizon past the tensor-memory kernel (16 QPs
    per SM no longer fit shared memory); 20 is the horizon the reference's comments quote for the controller (controllerMain.py:138,145)."""
    B = 21
    w = W.controller_batch(B, N, seed=200 + N)
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
    assert s.info()["variant"] == (6 if N <= 8 else 5)
    r = s.solve(w["x0"], **{k: w[k] for k in KEYS})
    _check_ctrl(track, N, w, r, range(B))
    s.close()


def test_planner_longest_masked_horizon(track):
    """Planner N = 63: the last horizon whose per-stage row classes fit the kernels' 64-bit stage masks."""
    N, B = 63, 5
    w = W.planner_batch(B, N, seed=4)
    s = lp.BatchSolver("planner", N, W.PLAN_DT, track=track, max_batch=B, **W.PLAN)
    assert s.info()["variant"] in (5, 7)
    keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], track, L_cf=W.PLAN["L_cf"])
    st = oracle.default_settings(polish=1)
    for b in range(B):
        o = oracle.plan_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], SS=w["SS"][b], u_prev=w["u_prev"][b],
                              u_old=w["u_old"][b], max_ey=float(w["max_ey"][b]), ey_lo=w["ey_lo"][b], ey_hi=w["ey_hi"][b])
        assert int(r.status[b]) == o["status"] and int(r.iters[b]) == o["iter"], (b, r.status[b], o["status"], r.iters[b], o["iter"])
        if o["status"] in (1, 2, -2):
            np.testing.assert_allclose(r.u_pred[b], o["uPred"], rtol=0, atol=1e-4)
    s.close()


def test_bad_problems_do_not_disturb_their_warp_neighbours(track):
    """Lap 0 with NaN / negative arc length (Curvature raises in the reference, utilities.py:44-46), an infeasible start
    (vx above max_vel with x0 pinned) and a NaN state, interleaved with good problems in the same 4-QP warp groups."""
    N, B = 8, 32
    w = W.controller_batch(B, N, seed=77)
    w["lap"][:] = 0
    w["x0"][:, 4] = np.abs(w["x0"][:, 4])
    w["x0"][1, 4] = np.nan        # schedule error
    w["x0"][6, 4] = -0.5          # schedule error (s < 0)
    w["x0"][9, 0] = 7.5           # infeasible: state bound vx <= 5 on the pinned x_0
    w["x0"][14, 1] = np.nan       # NaN dynamics
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
    r = s.solve(w["x0"], lap_all=0, **{k: w[k] for k in KEYS})
    assert int(r.status[1]) == -20 and int(r.status[6]) == -20
    assert int(r.status[9]) in (-3, 3)
    assert int(r.status[14]) not in (1, 2)
    good = [b for b in range(B) if b not in (1, 6, 9, 14)]
    _check_ctrl(track, N, w, r, good)
    s.close()


def test_fleet_with_the_long_controller_horizon():
    """BASELINE configs[0] 'also report at N = 20': the fleet tick with the 20-step horizon (all 20 rows of the warm-up
    guess of predicted_vectors_generation are used) against the oracle loop."""
    m = lp.Map("L_shape")
    B, T, N = 24, 30, 20
    sim0 = lp.fleet_start(B, seed=11, track_map=m)
    fleet = lp.ClosedLoopFleet(m, N=N, max_fleet=B)
    got = fleet.start(sim0).run(T).read()
    cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_PT["Q"], W.CTRL_PT["R"], W.CTRL_PT["dR"], m.PointAndTangent)
    ref = oracle.loop_state(sim0, N)
    oracle.loop_run(cfg, oracle.default_settings(polish=1), oracle.loop_cfg(half_width=m.halfWidth, slack=m.slack), ref, T, threads=8)
    np.testing.assert_array_equal(got["ctr"], ref["ctr"])
    np.testing.assert_array_equal(got["stat"][:, [0, 1, 3]], ref["stat"][:, [0, 1, 3]])
    np.testing.assert_allclose(got["sim"], ref["sim"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["cmd"], ref["cmd"], rtol=0, atol=1e-6)
    fleet.close()


def test_host_copy_modes_agree(track, monkeypatch):
    """The _host entry points return the same bits whether the kernel writes its results into the pinned arena
    (default), the results come back by a D2H copy (LPVMPC_ZERO_COPY_OUT=0), or the inputs are read from the arena too
    (LPVMPC_ZERO_COPY_IN=1); the modes are read when the handle is created."""
    N, B = 8, 300
    w = W.controller_batch(B, N, seed=17)
    outs = ("active_lo", "active_up", "y", "A_out", "states_out")
    res = []
    for zo, zi in (("1", "0"), ("0", "0"), ("1", "1")):
        monkeypatch.setenv("LPVMPC_ZERO_COPY_OUT", zo)
        monkeypatch.setenv("LPVMPC_ZERO_COPY_IN", zi)
        s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
        res.append(s.solve(w["x0"], extra_outputs=outs, **{k: w[k] for k in KEYS}))
        sch = s.schedule(x0=w["x0"], **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap")})
        res[-1]["sched_A"] = sch["A_out"]
        res[-1]["sched_err"] = sch["sched_err"]
        s.close()
    for r in res[1:]:
        for k in ("x_pred", "u_pred", "status", "iters", "obj", "pri_res", "dua_res", "sched_A", "sched_err") + outs:
            np.testing.assert_array_equal(r[k], res[0][k], err_msg=k)
    assert (res[0].status == 1).mean() > 0.9
