"""How many condensed-polish refinement passes does the device need to reproduce the oracle's polish decisions?
GPU passes = polish_refine_iter + kPolishExtraRefine (3); the oracle keeps OSQP's 3 on its full-KKT polish."""
import os, sys, time, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import oracle, lpvmpc_b200 as lp, torch
W = lp.workloads; track = lp.Map("L_shape").PointAndTangent
N, B = 8, 4096
w = W.controller_batch(B, N, seed=0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
st = oracle.default_settings(polish=1)
o = oracle.ctrl_batch(cfg, st, w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"], threads=os.cpu_count())
tin = {k: torch.as_tensor(w[k]).cuda() for k in keys}; tx0 = torch.as_tensor(w["x0"]).cuda()
for refine in (3, 2, 1, 0):
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, polish_refine_iter=refine, **W.CTRL_TT)
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1))
    for _ in range(5): s.solve(tx0, **tin)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.solve(tx0, **tin)
    e1.record(); torch.cuda.synchronize()
    print("passes", refine + 3, "status/iters eq", np.array_equal(r.status, o["status"]), np.array_equal(r.iters, o["iters"]), "polish ok", int((r.polish_status == 1).sum()),
          "d>1e-4:", int((d > 1e-4).sum()), "d>1e-6:", int((d > 1e-6).sum()), "d>1e-8:", int((d > 1e-8).sum()), "ms %.3f" % (e0.elapsed_time(e1) / 20))
    s.close()
