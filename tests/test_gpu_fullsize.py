"""Parity at BASELINE.json's FULL sizes: every QP of the bench workloads (configs[1], [2], [4]) against the CPU oracle's
batch loop on the same seeded inputs — status and ADMM iteration count identical for every problem, solutions of the
solved problems within 1e-4.

Documented exceptions, at most 0.2 % of a batch, each with the same status and iteration count and objectives within
1e-6 relative: (i) ill-conditioned long-horizon problems whose Euler roll-out has diverged (|x| ~ 1e3 at N = 100: 2 of
1,024), compared relative to |x|_inf; (ii) a handful of QPs (4 of 4,096 at configs[1]) where the polish accept / reject
decision differs — the device's condensed polish (3 extra refinement passes) reaches residuals of 1e-15 / 1e-11 and is accepted,
the oracle's full-KKT polish does not improve both residuals and is rejected, so the oracle keeps its ADMM iterate
(accurate to OSQP's eps = 1e-3).  For those the test requires: same status and iteration count, device residuals not
larger than the oracle's, objectives within 1e-6 relative, solutions within OSQP's own termination tolerance; and that they stay
below 0.2 % of the batch; (iii) 2 of the 16,384 planner QPs whose polish is marginal on both sides (accepted with
post-polish residuals of 1e-4 / 1e-3), differing by up to 6e-3 with objectives within 1.5e-4 relative."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

lp = pytest.importorskip("lpvmpc_b200")
W = lp.workloads
THREADS = os.cpu_count() or 1


def _track():
    return lp.Map("L_shape").PointAndTangent


@pytest.mark.parametrize("N,B,seed,steer", [(8, 4096, 0, 1.0), (100, 1024, 3, 0.2)])
def test_controller_bench_workloads_every_qp(N, B, seed, steer):
    track = _track()
    w = W.controller_batch(B, N, seed=seed, steer_scale=steer)
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
    r = s.solve(w["x0"], **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
    cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
    o = oracle.ctrl_batch(cfg, oracle.default_settings(polish=1), w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"],
                          threads=THREADS)
    np.testing.assert_array_equal(r.status, o["status"])
    np.testing.assert_array_equal(r.iters, o["iters"])
    ok = np.isin(o["status"], (1, 2, -2))
    assert ok.mean() > 0.99
    d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1))
    d[~ok] = 0.0
    scale = np.maximum(1.0, np.abs(np.nan_to_num(o["xPred"])).reshape(B, -1).max(1))   # N = 100 roll-outs reach |x| ~ 2e3
    odd = np.nonzero(d >= 1e-4 * scale)[0]
    assert len(odd) <= 0.002 * B, len(odd)
    st = oracle.default_settings(polish=1)
    for b in odd:   # see the module docstring
        oo = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                               curv_ref=w["curv_ref"][b], lap=int(w["lap"][b]), old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
        assert abs(r.obj[b] - oo["obj_val"]) <= 1e-6 * abs(oo["obj_val"]), b
        assert d[b] < 1e-2 * scale[b], (b, d[b])   # OSQP's own tolerance: eps_abs + eps_rel * |z|_inf
        if int(r.polish_status[b]) != oo["status_polish"]:
            assert int(r.polish_status[b]) == 1 and oo["status_polish"] == -1, (b, r.polish_status[b], oo["status_polish"])
            assert r.pri_res[b] <= oo["pri_res"] and r.dua_res[b] <= oo["dua_res"], b
    s.close()


def test_planner_bench_workload_every_qp():
    track = _track()
    N, B = 40, 16384
    w = W.planner_batch(B, N, seed=1)
    keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
    s = lp.BatchSolver("planner", N, W.PLAN_DT, track=track, max_batch=B, **W.PLAN)
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], track, L_cf=W.PLAN["L_cf"])
    o = oracle.plan_batch(cfg, oracle.default_settings(polish=1), w["x0"], w["SS"], w["u_prev"], w["u_old"], w["max_ey"], w["ey_lo"], w["ey_hi"],
                          threads=THREADS)
    np.testing.assert_array_equal(r.status, o["status"])
    np.testing.assert_array_equal(r.iters, o["iters"])
    ok = o["status"] == 1
    assert ok.mean() > 0.98
    d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1))
    d[~ok] = 0.0
    odd = np.nonzero(d >= 1e-4)[0]
    assert len(odd) <= 0.001 * B, len(odd)      # measured: 2 of 16,384 (10 above 1e-6)
    st = oracle.default_settings(polish=1)
    for b in odd:   # marginal polish (post-polish residuals ~1e-4 / 1e-3 on both sides): see the module docstring
        oo = oracle.plan_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], SS=w["SS"][b], u_prev=w["u_prev"][b], u_old=w["u_old"][b],
                               max_ey=float(w["max_ey"][b]), ey_lo=w["ey_lo"][b], ey_hi=w["ey_hi"][b])
        assert abs(r.obj[b] - oo["obj_val"]) <= 1e-3 * abs(oo["obj_val"]), (b, r.obj[b], oo["obj_val"])
        assert d[b] < 2e-2, (b, d[b])
    s.close()
