import os, sys, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import oracle, lpvmpc_b200 as lp
W = lp.workloads; track = lp.Map("L_shape").PointAndTangent
N, B = 8, 4096
w = W.controller_batch(B, N, seed=0)
s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
r = s.solve(w["x0"], **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
o = oracle.ctrl_batch(cfg, st, w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"], threads=os.cpu_count())
d = np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1)
bad = np.nonzero(d > 1e-6)[0]
print("n bad", len(bad), "polish ok share", (r.polish_status == 1).mean(), "polish fail", (r.polish_status == -1).sum())
for b in bad[:12]:
    oo = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b], curv_ref=w["curv_ref"][b],
                           lap=1, old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
    print(b, "d=%.2e" % d[b], "gpu polish", int(r.polish_status[b]), "oracle polish", oo["status_polish"], "iters", int(r.iters[b]), oo["iter"],
          "obj gpu %.9g oracle %.9g" % (r.obj[b], oo["obj_val"]), "pri/dua gpu %.2e %.2e" % (r.pri_res[b], r.dua_res[b]), "oracle %.2e %.2e" % (oo["pri_res"], oo["dua_res"]))
