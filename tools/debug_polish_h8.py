import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import lpvmpc_b200 as lp
import oracle
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = W.controller_batch(B, N, seed=0)
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
ref = [oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                         curv_ref=w["curv_ref"][b], lap=int(w["lap"][b]), old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1])) for b in range(B)]
for variant in (3, 5):
    for pol in (1,):
        s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT, polish=pol)
        r = s.solve(w["x0"], extra_outputs=("active_lo", "active_up", "y"), **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
        for b in range(B):
            o = ref[b]
            if int(r.polish_status[b]) == o['status_polish'] and int(r.iters[b]) == o['iter']: continue
            print("v", variant, "polish", pol, "b", b, "status", int(r.status[b]), o["status"], "iters", int(r.iters[b]), o["iter"], "pol", int(r.polish_status[b]), o["status_polish"],
                  "pri %.3e dua %.3e obj %.10e | oracle obj %.10e" % (r.pri_res[b], r.dua_res[b], r.obj[b], o["obj_val"]),
                  "dx %.2e du %.2e" % (np.abs(r.x_pred[b] - o["xPred"]).max(), np.abs(r.u_pred[b] - o["uPred"]).max()),
                  "nact", int((r.active_lo[b] | r.active_up[b]).sum()), int((o["active_lo"] | o["active_up"]).sum()), flush=True)
        s.close()
