"""Quick check of the H8 kernels on reduced planner / long-horizon batches: parity with the oracle + kernel time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import oracle
import lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "plan"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 888
if which == "plan":
    N = 40
    w = W.planner_batch_harvest(B, N, seed=1)
    keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
    s = lp.BatchSolver("planner", N, W.PLAN_DT, track=track, max_batch=B, **W.PLAN)
    cfg = oracle.make_cfg("planner", N, W.PLAN_DT, W.PLAN["Q"], W.PLAN["R"], W.PLAN["dR"], track, L_cf=W.PLAN["L_cf"])
    o = oracle.plan_batch(cfg, oracle.default_settings(polish=1), w["x0"], w["SS"], w["u_prev"], w["u_old"], w["max_ey"], w["ey_lo"], w["ey_hi"], threads=os.cpu_count())
else:
    N = int(which)
    w = W.controller_batch(B, N, seed=3, steer_scale=0.2)
    keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
    cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
    o = oracle.ctrl_batch(cfg, oracle.default_settings(polish=1), w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"], threads=os.cpu_count())
print("info", s.info())
if len(sys.argv) > 3 and sys.argv[3] == "fixed":   # time of a fixed number of plain ADMM steps (no checks, no polish)
    for iters in (100, 300):
        s.update_settings(max_iter=iters, check_termination=0, adaptive_rho=0, polish=0)
        tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"]).to(dev)
        for _ in range(2): s.solve(tx0, **tin)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            a.record(); s.solve(tx0, **tin); b.record()
        torch.cuda.synchronize()
        print("fixed %d iterations: kernel ms p50 %.4f" % (iters, np.percentile([a.elapsed_time(b) for a, b in ev], 50)))
    sys.exit(0)
r = s.solve(w["x0"], **{k: w[k] for k in keys})
if o is not None:
    ok = np.isin(o["status"], (1, 2, -2))
    d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1)); d[~ok] = 0
    print("status diff", int((r.status != o["status"]).sum()), "iter diff", int((r.iters != o["iters"]).sum()), "odd(>=1e-4)", np.nonzero(d >= 1e-4)[0][:10],
          "max d", float(np.nanmax(d)), "polish+", int((r.polish_status == 1).sum()), "iters mean", float(r.iters.mean()))
tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"]).to(dev)
for _ in range(2): s.solve(tx0, **tin)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for a, b in ev:
    a.record(); s.solve(tx0, **tin); b.record()
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in ev])
print("kernel ms p50 %.4f min %.4f" % (np.percentile(t, 50), t.min()))
