import os, sys, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import oracle, lpvmpc_b200 as lp
W = lp.workloads; track = lp.Map("L_shape").PointAndTangent
N, B = 100, 1024
w = W.controller_batch(B, N, seed=3, steer_scale=0.2)
s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
r = s.solve(w["x0"], **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
o = oracle.ctrl_batch(cfg, st, w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"], threads=os.cpu_count())
du = np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1); dx = np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1)
ok = np.isin(o["status"], (1, 2, -2)); du[~ok] = 0; dx[~ok] = 0
print("du>1e-4:", (du > 1e-4).sum(), "dx>1e-4:", (dx > 1e-4).sum(), "dx>1e-6", (dx > 1e-6).sum(), "max |x|", np.nanmax(np.abs(o["xPred"][ok])))
for b in np.nonzero(dx > 1e-4)[0][:10]:
    oo = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b], curv_ref=w["curv_ref"][b],
                           lap=1, old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
    k = np.unravel_index(np.abs(r.x_pred[b] - oo["xPred"]).argmax(), oo["xPred"].shape)
    print(b, "du %.2e dx %.2e at" % (du[b], dx[b]), k, "val %.6g" % oo["xPred"][k], "polish", int(r.polish_status[b]), oo["status_polish"], "iters", int(r.iters[b]),
          "res gpu %.1e %.1e oracle %.1e %.1e" % (r.pri_res[b], r.dua_res[b], oo["pri_res"], oo["dua_res"]), "obj rel %.1e" % (abs(r.obj[b] - oo["obj_val"]) / abs(oo["obj_val"])))
