import os, sys, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import oracle, lpvmpc_b200 as lp
W = lp.workloads; track = lp.Map("L_shape").PointAndTangent
N, B = 40, 16384
w = W.planner_batch(B, N, seed=1)
keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
s = lp.BatchSolver("planner", N, W.PLAN_DT, track=track, max_batch=B, **W.PLAN)
r = s.solve(w["x0"], **{k: w[k] for k in keys})
cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], track, L_cf=W.PLAN["L_cf"])
st = oracle.default_settings(polish=1)
o = oracle.plan_batch(cfg, st, w["x0"], w["SS"], w["u_prev"], w["u_old"], w["max_ey"], w["ey_lo"], w["ey_hi"], threads=os.cpu_count())
ok = o["status"] == 1
d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1)); d[~ok] = 0
print("status eq", np.array_equal(r.status, o["status"]), "iters eq", np.array_equal(r.iters, o["iters"]))
print("d>1e-4:", (d > 1e-4).sum(), "d>1e-6:", (d > 1e-6).sum(), "gpu polish ok", (r.polish_status == 1).sum(), "fail", (r.polish_status == -1).sum())
odd = np.nonzero(d > 1e-4)[0]
same = diff = 0; worst_same = 0.0
for b in odd:
    oo = oracle.plan_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], SS=w["SS"][b], u_prev=w["u_prev"][b], u_old=w["u_old"][b],
                           max_ey=float(w["max_ey"][b]), ey_lo=w["ey_lo"][b], ey_hi=w["ey_hi"][b])
    if int(r.polish_status[b]) == oo["status_polish"]:
        same += 1; worst_same = max(worst_same, d[b])
        if same <= 5: print("SAME", b, d[b], r.polish_status[b], "res gpu %.1e %.1e or %.1e %.1e" % (r.pri_res[b], r.dua_res[b], oo["pri_res"], oo["dua_res"]), "objrel %.1e" % (abs(r.obj[b]-oo["obj_val"])/abs(oo["obj_val"])))
    else:
        diff += 1
        if diff <= 5: print("DIFF", b, d[b], int(r.polish_status[b]), oo["status_polish"], "res gpu %.1e %.1e or %.1e %.1e" % (r.pri_res[b], r.dua_res[b], oo["pri_res"], oo["dua_res"]), "objrel %.1e" % (abs(r.obj[b]-oo["obj_val"])/abs(oo["obj_val"])))
print("same-decision odd:", same, "worst", worst_same, "different-decision odd:", diff)
