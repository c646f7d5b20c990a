"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): one launch of every kernel the library
ships — solve kernels variant 1 (generic), 5 (H8, controller and planner, 4 / 1 QPs per warp), 6 (H8T, tensor memory),
7 (H8S, TMA-streamed factor), 8 (H16T: CTA rounds), the twisted H8 kernels with helper warps (v5n100, p5) and the plain one (p5n39), the visiting-order kernel, the stand-alone scheduling kernels,
the fleet / planner loop kernels, the reference and hand-off kernels.  Results are checked against the oracle so a
sanitizer-clean run is also a correct one.

    compute-sanitizer --tool racecheck python tools/sanitize_subset.py [names...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import lpvmpc_b200 as lp  # noqa: E402

W = lp.workloads
M = lp.Map("L_shape")
TRACK = M.PointAndTangent
KEYS = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
SETTINGS = dict(max_iter=150)   # bounded: the sanitizers slow a kernel down 10-100x


def ctrl(variant, N=8, B=10):
    w = W.controller_batch(B, N, seed=7, steer_scale=(0.2 if N >= 50 else 1.0))
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=TRACK, max_batch=B, variant=variant, **W.CTRL_TT, **SETTINGS)
    r = s.solve(w["x0"], **{k: w[k] for k in KEYS})
    cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], TRACK)
    st = oracle.default_settings(polish=1, **SETTINGS)
    for b in range(B):
        o = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                              curv_ref=w["curv_ref"][b], lap=1, old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
        assert int(r.status[b]) == o["status"] and int(r.iters[b]) == o["iter"], (variant, N, b, int(r.status[b]), o["status"], int(r.iters[b]), o["iter"])
    info = s.info()
    s.close()
    return "variant %d N %d: ok (%d launches)" % (info["variant"], N, info["kernel_launches"])


def plan(variant, B=5, N=40):
    w = W.planner_batch(B, N, seed=7)
    keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
    s = lp.BatchSolver("planner", N, W.PLAN_DT, track=TRACK, max_batch=B, variant=variant, **W.PLAN, **SETTINGS)
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], TRACK, L_cf=W.PLAN["L_cf"])
    st = oracle.default_settings(polish=1, **SETTINGS)
    for b in range(B):
        o = oracle.plan_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], SS=w["SS"][b], u_prev=w["u_prev"][b], u_old=w["u_old"][b],
                              max_ey=float(w["max_ey"][b]), ey_lo=w["ey_lo"][b], ey_hi=w["ey_hi"][b])
        assert int(r.status[b]) == o["status"] and int(r.iters[b]) == o["iter"], (variant, b)
    info = s.info()
    s.close()
    return "planner variant %d: ok" % info["variant"]


def schedule():
    N, B = 8, 70
    w = W.controller_batch(B, N, seed=3)
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=TRACK, max_batch=B, **W.CTRL_TT)
    r = s.schedule(x0=w["x0"], u_prev=w["u_prev"], vel_ref=w["vel_ref"], curv_ref=w["curv_ref"], lap=w["lap"])
    s.schedule(sched_mode=lp.SCHED_ESTIMATE, traj=r.states_out.copy(), u_prev=w["u_prev"])
    s.close()
    wp = W.planner_batch(33, 40, seed=3)
    sp = lp.BatchSolver("planner", 40, W.PLAN_DT, track=TRACK, max_batch=33, **W.PLAN)
    sp.schedule(x0=wp["x0"], SS=wp["SS"], u_prev=wp["u_prev"])
    sp.close()
    return "schedule kernels: ok"


def fleet():
    sim0 = lp.fleet_start(24, seed=3, track_map=M)
    f = lp.ClosedLoopFleet(M, N=8, max_fleet=24, **SETTINGS)
    got = f.start(sim0).run(11).read()
    assert (got["ctr"][:, 5] == 0).all()
    f.close()
    return "fleet loop: ok"


def planfleet():
    rng = np.random.default_rng(5)
    B = 4
    x0 = np.stack([rng.uniform(1.0, 2.0, B), rng.normal(0, 0.01, B), rng.normal(0, 0.05, B), rng.normal(0, 0.02, B), rng.normal(0, 0.02, B)], axis=1)
    f = lp.PlannerFleet(M, N=40, max_fleet=B, max_ey=0.2, **SETTINGS)
    f.start(x0, rng.uniform(0.0, 18.0, B))
    f.run(2)
    out = f.read()
    refs = f.references(out["x_pred"], out["SS"], np.zeros((B, 3)))
    assert refs is not None
    f.close()
    return "planner loop + references: ok (%d ticks)" % int(out["ctr"][:, 0].sum())


def aux():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "aux.npz")))
    s = lp.BatchSolver("controller", 8, W.CTRL_DT, track=TRACK, max_batch=64, **W.CTRL_TT)
    r = lp.anfis_abc(s, g["an_sched"], g["an_A_tab"], g["an_B_tab"], g["an_C_tab"], g["an_bell"])
    assert np.allclose(r["A"], g["an_A"], rtol=1e-12, atol=1e-14)
    e = lp.observer_step(s, g["ob_est0"], g["ob_y"][0], g["ob_u"][0], g["ob_lim_ls"], g["ob_gains_ls"], g["ob_lim_hs"], g["ob_gains_hs"], g["ob_C"],
                         dt=float(g["ob_dt"]), use_estimate=0)
    assert np.allclose(e, g["ob_est"][0], rtol=1e-11, atol=1e-12)
    s.close()
    return "TS-fuzzy blend + observer kernels: ok"


CASES = {
    "v1": lambda: ctrl(1), "v5": lambda: ctrl(5), "v6": lambda: ctrl(6), "v7": lambda: ctrl(7), "v8": lambda: ctrl(8, B=37),
    "p5n39": lambda: plan(5, B=3, N=39), "aux": aux,
    "v5n20": lambda: ctrl(5, N=20, B=6), "v5n100": lambda: ctrl(5, N=100, B=2), "v7n160": lambda: ctrl(7, N=160, B=2),
    "p1": lambda: plan(1, B=2), "p5": lambda: plan(5), "p7": lambda: plan(7, B=3),
    "schedule": schedule, "fleet": fleet, "planfleet": planfleet,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        print(n, CASES[n]())
    print("sanitize subset done")
