import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import oracle, lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, 512
w = W.controller_batch(B, N, seed=0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
O = [oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                       curv_ref=w["curv_ref"][b], lap=1, old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1])) for b in range(B)]
opol = np.array([o["status_polish"] for o in O]); ox = np.array([o["xPred"] for o in O]); ou = np.array([o["uPred"] for o in O])
odua = np.array([o["dua_res"] for o in O]); opri = np.array([o["pri_res"] for o in O])
for refine in (3, 5, 8, 12, 20):
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, polish_refine_iter=refine, **W.CTRL_TT)
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    mism = (r.polish_status != opol)
    dx = np.abs(r.x_pred - ox).reshape(B, -1).max(1); du = np.abs(r.u_pred - ou).reshape(B, -1).max(1)
    print("refine", refine, "polish mismatches", int(mism.sum()), "max dx %.2e du %.2e" % (dx.max(), du.max()),
          "median gpu dua %.2e pri %.2e | oracle dua %.2e pri %.2e" % (np.median(r.dua_res), np.median(r.pri_res), np.median(odua), np.median(opri)),
          "p99 gpu dua %.2e" % np.percentile(r.dua_res, 99), "worst-b", int(dx.argmax()))
    s.close()
