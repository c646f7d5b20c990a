"""Generate tests/golden/closed_loop.npz from the REFERENCE'S OWN Python (imported headless from /root/reference).

Run in the build container only:

    python tests/golden/make_golden_loop.py

Recorded (all produced by reference code):
  * Simulator.f             (vehicleSimulator.py:164-199; the class body is exec'd from the reference file after
                             ``expandtabs`` — the file mixes tabs and spaces — with rospy stubbed)
  * Map.getLocalPosition    (Utilities/trackInitialization.py:283-383) on on-track, off-track and vertex points
  * Map.getGlobalPosition   (Utilities/trackInitialization.py:205-260)
  * predicted_vectors_generation (controllerMain.py:510-553; the function is exec'd from the reference file)
  * a closed loop of the reference's controller class + Map + Simulator: the controller main loop
    (controllerMain.py:177-192, 289-298, 310-331, 381-383; it cannot be imported — ROS messages, matplotlib, hard-coded
    home paths) is restated below line by line; the QP is solved at the ``osqp`` stub seam by the oracle's OSQP
    restatement (the reference's solver is the absent PyPI package), so the *loop* is pinned, the solve is not.
"""
import os
import re
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

SIM_PARAMS = {"simulator/c_f": 0.8, "simulator/B": 6.0, "simulator/C": 1.6, "simulator/mu": 0.05,
              "simulator/init_vx": 0.2, "simulator/dt": 0.005}  # MAIN_LAUNCH.launch:60-77


def load_simulator(ns):
    """exec ``class Simulator`` from the reference's vehicleSimulator.py."""
    path = os.path.join(refload.REF_SRC, "vehicleSimulator.py")
    src = open(path).read().expandtabs(4)
    m = re.search(r"^class Simulator\(object\):.*?(?=^class )", src, re.S | re.M)
    body = refload._PRINT_RE.sub(r"\1print(\2)", m.group(0))
    ns.rospy._params.update(SIM_PARAMS)
    ns.rospy.Rate = lambda hz: None
    ns.rospy.get_rostime = lambda: None
    g = {"rospy": ns.rospy}
    exec("from numpy import tan, arctan, cos, sin, pi", g)
    exec(compile(body, path, "exec"), g)
    return g["Simulator"]


def load_guess():
    path = os.path.join(refload.REF_SRC, "controllerMain.py")
    src = open(path).read()
    m = re.search(r"^def predicted_vectors_generation\(.*?^    return xx, uu", src, re.S | re.M)
    g = {"np": np}
    exec(compile(m.group(0), path, "exec"), g)
    return g["predicted_vectors_generation"]


def sim_cases(Simulator):
    rng = np.random.default_rng(7)
    n = 256
    st = np.stack([rng.uniform(-3, 6, n), rng.uniform(-3, 6, n), rng.uniform(-7, 7, n), rng.uniform(0.0, 3.0, n),
                   rng.uniform(-0.5, 0.5, n), rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], axis=1)
    st[:32, 3] = rng.uniform(0.0, 0.25, 32)  # around the |vx| > 0.2 switch
    st[32, 3] = 0.2
    u = np.stack([rng.uniform(-1, 4, n), rng.uniform(-0.249, 0.249, n)], axis=1)  # [motor, servo]
    out = np.zeros_like(st)
    sim = Simulator()
    for i in range(n):
        sim.x, sim.y, sim.yaw, sim.vx, sim.vy, sim.psiDot, sim.ax, sim.ay = st[i]
        sim.f(u[i])
        out[i] = [sim.x, sim.y, sim.yaw, sim.vx, sim.vy, sim.psiDot, sim.ax, sim.ay]
    return {"sim_state": st, "sim_u": u, "sim_next": out}


def position_cases(track_map):
    rng = np.random.default_rng(8)
    L = track_map.TrackLength
    s = np.r_[rng.uniform(0, L, 300), np.linspace(0.0, L, 41)[:-1]]
    ey = np.r_[rng.uniform(-0.3, 0.3, 300), np.zeros(40)]
    gp = np.array([track_map.getGlobalPosition(si, ei) for si, ei in zip(s, ey)], dtype=np.float64).reshape(-1, 3)
    dpsi = np.r_[rng.uniform(-0.6, 0.6, 300), np.zeros(40)]
    pts = np.c_[gp[:, 0], gp[:, 1], gp[:, 2] + dpsi]
    # yaw several turns away (np.unwrap path), off-track points, the segment vertices themselves
    far = pts[:40].copy(); far[:, 2] += 2 * np.pi * rng.integers(-3, 4, 40)
    off = np.c_[rng.uniform(-4, 8, 60), rng.uniform(-4, 8, 60), rng.uniform(-3, 3, 60)]
    pt = track_map.PointAndTangent
    vert = np.c_[pt[:, 0], pt[:, 1], pt[:, 2] + 0.1]
    pts = np.r_[pts, far, off, vert]
    loc = np.array([track_map.getLocalPosition(*p) for p in pts], dtype=np.float64)
    return {"gp_s": s, "gp_ey": ey, "gp_out": gp, "lp_in": pts, "lp_out": loc,
            "halfWidth": np.array(track_map.halfWidth), "slack": np.array(track_map.slack)}


def closed_loop(ns, Simulator, guess, track_map, swap, ticks, N=8, substeps=7):
    """controllerMain.py main loop, lap 0 (path tracking tune :139-141), true-state feedback."""
    Q = np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0])
    R = 0.5 * 0.5 * np.diag([1.0, 1.0])
    dR = 1.5 * 25 * np.array([1.3, 1.0])
    dt = 1.0 / 30.0
    Controller = ns.PathFollowingLPV_MPC(Q, R, dR, N, 1.0, dt, track_map, "OSQP", 0, 0)
    st = oracle.default_settings(polish=1)
    statuses = []

    def backend(qp):
        r = oracle.osqp_solve(qp.P, qp.q, qp.A, qp.l, qp.u, settings=st)
        statuses.append((r["status"], r["iter"]))
        return r["x"], r["status"]

    ns.OSQPSeam.backend = staticmethod(backend)
    sim = Simulator()
    servo = motor = 0.0
    first_it = 1
    rec = {k: [] for k in ("sim", "local", "cmd", "upred", "xpred")}
    uApplied = np.array([0.0, 0.0])
    try:
        for _ in range(ticks):
            rec["sim"].append([sim.x, sim.y, sim.yaw, sim.vx, sim.vy, sim.psiDot, sim.ax, sim.ay])
            GlobalState = np.array([sim.vx, sim.vy, sim.psiDot, sim.x, sim.y, sim.yaw])  # estimator bypassed
            LocalState = GlobalState.copy()
            if LocalState[0] < 0.01:
                LocalState[0] = 0.01
            if swap:   # as written, controllerMain.py:188
                LocalState[4], LocalState[3], LocalState[5], inside = track_map.getLocalPosition(GlobalState[3], GlobalState[4], GlobalState[5])
            else:      # what the comment on :187 intends (s, epsi, ey)
                LocalState[4], LocalState[5], LocalState[3], inside = track_map.getLocalPosition(GlobalState[3], GlobalState[4], GlobalState[5])
            rec["local"].append(LocalState.copy())
            uApplied = np.array([servo, motor])
            Controller.OldSteering.append(servo); Controller.OldAccelera.append(motor)
            Controller.OldSteering.pop(0); Controller.OldAccelera.pop(0)
            if first_it < 10:
                vel_ref = np.ones(N)
                xx, uu = guess(N, LocalState, uApplied, dt)
                Controller.solve(LocalState[0:6], xx, uu, False, vel_ref, 0, 0, 0, first_it)
                first_it = first_it + 1
            else:
                vel_ref = np.ones(N + 1)
                curv_ref = np.zeros(N)
                pred, A_L, B_L, C_L = Controller.LPVPrediction(LocalState[0:6], Controller.uPred, vel_ref, curv_ref, 60, 0)
                Controller.solve(pred[0, :], pred, Controller.uPred, False, vel_ref, A_L, B_L, C_L, first_it)
            servo = float(Controller.uPred[0, 0])
            motor = float(Controller.uPred[0, 1])
            rec["cmd"].append([servo, motor])
            rec["upred"].append(np.array(Controller.uPred))
            rec["xpred"].append(np.array(Controller.xPred))
            for _k in range(substeps):
                sim.f([motor, servo])
    finally:
        ns.OSQPSeam.backend = None
    p = "loop%d_" % swap
    out = {p + k: np.array(v) for k, v in rec.items()}
    out[p + "status"] = np.array(statuses)
    out[p + "final_sim"] = np.array([sim.x, sim.y, sim.yaw, sim.vx, sim.vy, sim.psiDot, sim.ax, sim.ay])
    return out


def main():
    ns = refload.load()
    Simulator = load_simulator(ns)
    guess = load_guess()
    m = ns.Map()
    out = {}
    out.update(sim_cases(Simulator))
    out.update(position_cases(m))
    loc = np.array([0.3, 0.01, -0.02, 0.03, 2.5, -0.04])
    xx, uu = guess(8, loc, np.zeros(2), 1 / 30.0)
    out.update({"guess_local": loc, "guess_xx": xx, "guess_uu": uu})
    for swap in (1, 0):
        out.update(closed_loop(ns, Simulator, guess, m, swap, ticks=60))
    np.savez_compressed(os.path.join(OUT, "closed_loop.npz"), **out)
    print("closed_loop.npz", os.path.getsize(os.path.join(OUT, "closed_loop.npz")), "bytes")
    for swap in (1, 0):
        p = "loop%d_" % swap
        print("swap", swap, "final local", out[p + "local"][-1], "status set", sorted(set(out[p + "status"][:, 0].tolist())))


if __name__ == "__main__":
    main()
