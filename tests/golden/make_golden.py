"""Generate tests/golden/*.npz from the REFERENCE'S OWN Python, imported headless from /root/reference.

Run in the build container only (the reference checkout does not exist on the GPU box):

    python tests/golden/make_golden.py

What is recorded is produced by reference code, not by this repo:
  * Map().PointAndTangent for the L_shape and oval tracks     (trackInitialization.py:13-202)
  * Curvature(s) samples                                       (utilities.py:31-50)
  * controller LPVPrediction / _EstimateABC outputs            (PathFollowingLPVMPC.py:166-258,732-809)
  * planner LPVPrediction / _EstimateABC outputs               (LPV_MPC_Planner.py:242-320,519-591)
  * the exact (P, q, A, l, u, settings) each solve() hands to osqp.OSQP().setup, captured at the stub seam
    (PathFollowingLPVMPC.py:302-313, LPV_MPC_Planner.py:204-205), with P reduced to its upper triangle
    the way osqp's python wrapper does.
The OSQP solve itself has no reference-side golden (the package is absent: "parity unpinned").
"""
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# controllerMain.py:139-141 (path tracking) and :146-148 (trajectory tracking)
CTRL_TUNES = {
    "pt": (np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0]), 0.5 * 0.5 * np.diag([1.0, 1.0]), 1.5 * 25 * np.array([1.3, 1.0])),
    "tt": (np.diag([400.0, 1.0, 1.0, 20.0, 0.0, 1100.0]), 0.0 * np.diag([1.0, 1.0]), np.array([100.0, 45.0])),
}
# plannerMain.py:96-99
PLAN_Q = -np.diag([-0.000000000000088, -9.703658572659423, -0.5, 0.000000000213635, -0.153591566469547])
PLAN_L = -np.array([1.00702414775175, 0.187661946033823, -0.0, 0.0, -0.0329493219494661])
PLAN_R = np.diag([0.8, 0.0])
PLAN_DR = np.array([6.0, 6.0])


def qp_arrays(qp, prefix):
    P = sparse.triu(qp.P, format="csc")
    P.sort_indices()
    A = qp.A.tocsc()
    A.sort_indices()
    return {prefix + "P_data": P.data, prefix + "P_indices": P.indices, prefix + "P_indptr": P.indptr,
            prefix + "A_data": A.data, prefix + "A_indices": A.indices, prefix + "A_indptr": A.indptr,
            prefix + "q": qp.q, prefix + "l": qp.l, prefix + "u": qp.u,
            prefix + "shape": np.array([A.shape[0], A.shape[1]]),
            prefix + "settings": np.array(sorted("%s=%r" % kv for kv in qp.settings.items()))}


def controller_cases(ns, track_map):
    out = {}
    rng = np.random.default_rng(20240)
    idx = 0
    for tune, (Q, R, dR) in CTRL_TUNES.items():
        for N in (8, 20):
            for lap in (0, 1):
                for delay in ((0, 2) if (N == 8 and tune == "pt") else (0,)):
                    C = ns.PathFollowingLPV_MPC(Q, R, dR, N, 1.0, 1.0 / 30.0, track_map, "OSQP", delay, 0)
                    x = np.array([rng.uniform(0.5, 3.0), rng.uniform(-0.2, 0.2), rng.uniform(-1, 1),
                                  rng.uniform(-0.2, 0.2), rng.uniform(0, 19.2), rng.uniform(-0.2, 0.2)])
                    u = np.c_[np.clip(rng.uniform(-0.2, 0.2) + 0.01 * np.cumsum(rng.standard_normal(N)), -0.249, 0.249),
                              np.full(N, rng.uniform(-0.5, 1.5))]
                    vel_ref = np.minimum(rng.uniform(0.8, 3.0) + 0.05 * np.arange(N + 1), 5.0)
                    curv_ref = rng.choice([0.0, 0.6981317, -0.6981317], size=N)
                    C.OldSteering = [float(v) for v in rng.uniform(-0.2, 0.2, size=1 + delay)]
                    C.OldAccelera = [float(rng.uniform(-0.5, 1.0))]
                    states, A, B, Cc = C.LPVPrediction(x, u, vel_ref, curv_ref, 60, lap)
                    x0 = states[0] if lap == 0 else x
                    ns.OSQPSeam.log.clear()
                    C.solve(x0, states, u, False, vel_ref, A, B, Cc, 20)
                    p = "c%d_" % idx
                    out.update({p + "tune": np.array(tune), p + "N": np.array(N), p + "lap": np.array(lap),
                                p + "delay": np.array(delay), p + "Q": Q, p + "R": R, p + "dR": dR,
                                p + "x": x, p + "u": u, p + "vel_ref": vel_ref, p + "curv_ref": curv_ref,
                                p + "old_steering": np.array(C.OldSteering), p + "old_accel": np.array(C.OldAccelera[0]),
                                p + "states": states, p + "A": np.array(A), p + "B": np.array(B), p + "C": np.array(Cc)[:, :, 0],
                                p + "x0": np.array(x0)})
                    out.update(qp_arrays(ns.OSQPSeam.log[-1], p + "qp_"))
                    # warm-up path (_EstimateABC from a given trajectory, first_it < 10): lap-0 geometry
                    if delay == 0:
                        traj = np.c_[states, np.zeros(N)][:, :6]
                        traj[:, 0] = np.maximum(traj[:, 0], 0.3)
                        traj[:, 4] = np.mod(traj[:, 4], 19.0)
                        ns.OSQPSeam.log.clear()
                        C.solve(x, traj, u, False, vel_ref[:N], 0, 0, 0, 3)
                        out.update({p + "wu_traj": traj, p + "wu_A": np.array(C.A), p + "wu_B": np.array(C.B)})
                        out.update(qp_arrays(ns.OSQPSeam.log[-1], p + "wu_qp_"))
                    idx += 1
    out["n_cases"] = np.array(idx)
    return out


def planner_cases(ns, track_map):
    out = {}
    rng = np.random.default_rng(20241)
    idx = 0
    for N in (40, 12):
        for max_ey in (0.3, 0.2):
            P = ns.LPV_MPC_Planner(PLAN_Q, PLAN_R, PLAN_DR, PLAN_L, N, 1.0 / 20.0, track_map, "OSQP")
            x = np.array([rng.uniform(1.0, 2.5), rng.uniform(-0.05, 0.05), rng.uniform(-0.3, 0.3),
                          rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1)])
            s0 = rng.uniform(0, 19.0)
            SS = s0 + np.cumsum(np.r_[0.0, np.full(N, x[0] / 20.0)])
            u = np.c_[np.clip(0.05 * np.cumsum(rng.standard_normal(N)) * 0.1, -0.249, 0.249), np.full(N, rng.uniform(-0.2, 0.8))]
            states, A, B, Cc = P.LPVPrediction(x, SS, u)
            ns.OSQPSeam.log.clear()
            P.solve(x, 0, 0, A, B, Cc, 5, max_ey)
            p = "p%d_" % idx
            out.update({p + "N": np.array(N), p + "max_ey": np.array(max_ey), p + "x": x, p + "SS": SS, p + "u": u,
                        p + "states": states, p + "A": np.array(A), p + "B": np.array(B)})
            out.update(qp_arrays(ns.OSQPSeam.log[-1], p + "qp_"))
            # warm-up path: trajectory [vx vy w ey epsi s] + steering vector (plannerMain.py:465-505)
            traj = np.c_[np.r_[[x], states][:N, :], SS[:N]]
            traj[:, 0] = np.maximum(traj[:, 0], 0.5)
            uu = u[:, 0].copy()
            ns.OSQPSeam.log.clear()
            P.solve(x, traj, uu, 0, 0, 0, 1, max_ey)
            out.update({p + "wu_traj": traj, p + "wu_uu": uu, p + "wu_A": np.array(P.A), p + "wu_B": np.array(P.B)})
            out.update(qp_arrays(ns.OSQPSeam.log[-1], p + "wu_qp_"))
            idx += 1
    out["n_cases"] = np.array(idx)
    return out


def track_cases(ns):
    out = {}
    m = ns.Map()
    out["L_shape_PointAndTangent"] = m.PointAndTangent
    out["L_shape_TrackLength"] = np.array(m.TrackLength)
    out["L_shape_halfWidth"] = np.array(m.halfWidth)
    mo = ns.Map(1)
    out["oval_PointAndTangent"] = mo.PointAndTangent
    out["oval_TrackLength"] = np.array(mo.TrackLength)
    s = np.r_[np.linspace(0.0, 3 * m.TrackLength, 301)[:-1], m.PointAndTangent[:, 3], m.PointAndTangent[:, 3] + 1e-12,
              m.PointAndTangent[1:, 3] - 1e-12, m.TrackLength - 1e-9, m.TrackLength + 1e-9]
    s = s[np.mod(s, m.TrackLength) != 0.0] if False else s
    kap = []
    ok = []
    for si in s:
        try:
            kap.append(ns.Curvature(si, m.PointAndTangent))
            ok.append(1)
        except TypeError:
            kap.append(np.nan)
            ok.append(0)
    out["curv_s"] = s
    out["curv_kappa"] = np.array(kap)
    out["curv_ok"] = np.array(ok)
    return out, m


def main():
    ns = refload.load()
    tr, m = track_cases(ns)
    np.savez_compressed(os.path.join(OUT, "track.npz"), **tr)
    np.savez_compressed(os.path.join(OUT, "controller.npz"), **controller_cases(ns, m))
    np.savez_compressed(os.path.join(OUT, "planner.npz"), **planner_cases(ns, m))
    for f in ("track.npz", "controller.npz", "planner.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
