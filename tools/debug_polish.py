import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import oracle, lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, 256
w = W.controller_batch(B, N, seed=0)
s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
r = s.solve(w["x0"], extra_outputs=("active_lo", "active_up", "xs", "zs", "ys"), **{k: w[k] for k in keys})
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
bad = 0
for b in range(B):
    o = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                          curv_ref=w["curv_ref"][b], lap=1, old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
    if int(r.polish_status[b]) != o["status_polish"] or b < 3:
        bad += 1
        print("b", b, "gpu pol", r.polish_status[b], "ora pol", o["status_polish"], "iters", r.iters[b], o["iter"],
              "gpu pri/dua %.3e %.3e" % (r.pri_res[b], r.dua_res[b]), "ora pri/dua %.3e %.3e" % (o["pri_res"], o["dua_res"]),
              "obj", r.obj[b], o["obj_val"], "dx", np.abs(r.x_pred[b] - o["xPred"]).max(), "du", np.abs(r.u_pred[b] - o["uPred"]).max(),
              "iter-rel", np.abs(r.xs[b]-o["xs"]).max()/np.abs(o["xs"]).max())
print("mismatching polish status:", bad - 3)
