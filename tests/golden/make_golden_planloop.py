"""Generate tests/golden/planner_loop.npz from the REFERENCE'S OWN Python (imported headless from /root/reference).

    python tests/golden/make_golden_planloop.py

Recorded: the planner main loop (plannerMain.py:128-224, restated below line by line because the file imports ROS
messages and matplotlib) driving the reference's LPV_MPC_Planner class, Curvature and predicted_vectors_generation
(exec'd from plannerMain.py:465-505), with the QP solved at the osqp stub seam by the oracle's OSQP restatement:
per tick xPred, uPred, SS, status and iteration count.  Testing-mode start [1, 0, 0, 0, 0] at s = 0
(plannerMain.py:145-146) and a second start state.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
Q = -np.diag([-0.000000000000088, -9.703658572659423, -0.5, 0.000000000213635, -0.153591566469547])   # plannerMain.py:96
L = -np.array([1.00702414775175, 0.187661946033823, -0.0, 0.0, -0.0329493219494661])
R = np.diag([0.8, 0.0])
dR = np.array([6.0, 6.0])


def load_guess():
    path = os.path.join(refload.REF_SRC, "plannerMain.py")
    src = open(path).read()
    m = re.search(r"^def predicted_vectors_generation\(.*?^    return xx, uu", src, re.S | re.M)
    g = {"np": np, "hstack": np.hstack}
    exec(compile(m.group(0), path, "exec"), g)
    return g["predicted_vectors_generation"]


def planner_loop(ns, guess, track_map, x0, ticks, N=40, dt=1.0 / 20.0, HW=0.2):
    Planner = ns.LPV_MPC_Planner(Q, R, dR, L, N, dt, track_map, "OSQP")
    st = oracle.default_settings(polish=1)
    log = []

    def backend(qp):
        r = oracle.osqp_solve(qp.P, qp.q, qp.A, qp.l, qp.u, settings=st)
        log.append((r["status"], r["iter"]))
        return r["x"], r["status"]

    ns.OSQPSeam.backend = staticmethod(backend)
    SS = np.zeros(N + 1)
    first_it = 1
    rec = {k: [] for k in ("xpred", "upred", "SS")}
    try:
        for _ in range(ticks):
            if first_it == 1:
                xx, uu = guess(N, np.array(x0, dtype=float), 0.2, dt)
                Planner.solve(np.array(x0, dtype=float), xx, uu.ravel(), 0, 0, 0, first_it, HW)   # numpy 2: (Hp,1) -> ravel, SURVEY 8c
                first_it += 1
            else:
                LPV_X_Pred, A_L, B_L, C_L = Planner.LPVPrediction(Planner.xPred[1, :], SS[:], Planner.uPred)
                Planner.solve(Planner.xPred[1, :], 0, 0, A_L, B_L, C_L, first_it, HW)
            Planner.OldSteering.append(Planner.uPred[0, 0])
            Planner.OldAccelera.append(Planner.uPred[0, 1])
            for j in range(0, N):
                curv = ns.Curvature(SS[j], track_map.PointAndTangent)
                SS[j + 1] = (SS[j] + ((Planner.xPred[j, 0] * np.cos(Planner.xPred[j, 4])
                             - Planner.xPred[j, 1] * np.sin(Planner.xPred[j, 4])) / (1 - Planner.xPred[j, 3] * curv)) * dt)
            SS[0] = SS[1]
            rec["xpred"].append(np.array(Planner.xPred)); rec["upred"].append(np.array(Planner.uPred)); rec["SS"].append(SS.copy())
    finally:
        ns.OSQPSeam.backend = None
    out = {k: np.array(v) for k, v in rec.items()}
    out["status"] = np.array(log)
    return out


def main():
    ns = refload.load()
    guess = load_guess()
    m = ns.Map()
    out = {}
    xx, uu = guess(40, np.array([1.3, 0.02, -0.1, 0.05, 0.03]), 0.2, 0.05)
    out.update({"guess_x0": np.array([1.3, 0.02, -0.1, 0.05, 0.03]), "guess_xx": xx})
    for i, x0 in enumerate(([1.0, 0.0, 0.0, 0.0, 0.0], [1.6, 0.01, 0.05, -0.03, 0.02])):
        r = planner_loop(ns, guess, m, x0, ticks=25)
        out.update({"loop%d_%s" % (i, k): v for k, v in r.items()})
        out["loop%d_x0" % i] = np.array(x0)
        print("loop", i, "statuses", sorted(set(r["status"][:, 0].tolist())), "iters", r["status"][:, 1].tolist()[:8], "SS end", r["SS"][-1][:3], "vx", r["xpred"][-1][0, 0])
    np.savez_compressed(os.path.join(OUT, "planner_loop.npz"), **out)
    print("planner_loop.npz", os.path.getsize(os.path.join(OUT, "planner_loop.npz")), "bytes")


if __name__ == "__main__":
    main()
