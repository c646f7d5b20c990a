"""Quick A/B of kernel variants on ctrl4096: parity with the oracle on every QP + kernel time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import oracle
import lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, 4096
w = W.controller_batch(B, N, seed=0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
o = oracle.ctrl_batch(cfg, oracle.default_settings(polish=1), w["x0"], w["u_prev"], w["vel_ref"], w["curv_ref"], w["lap"], w["u_old"], threads=os.cpu_count())
dev = torch.device("cuda", 0)
for variant in [int(v) for v in (sys.argv[1:] or ["6", "8"])]:
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT)
    r = s.solve(w["x0"], **{k: w[k] for k in keys})
    ok = np.isin(o["status"], (1, 2, -2))
    d = np.maximum(np.abs(r.u_pred - o["uPred"]).reshape(B, -1).max(1), np.abs(r.x_pred - o["xPred"]).reshape(B, -1).max(1)); d[~ok] = 0
    print("variant", variant, "status diff", int((r.status != o["status"]).sum()), "iter diff", int((r.iters != o["iters"]).sum()),
          "odd(>=1e-4)", np.nonzero(d >= 1e-4)[0][:10], "max d", float(np.nanmax(d)), "polish+", int((r.polish_status == 1).sum()))
    tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"]).to(dev)
    for _ in range(5): s.solve(tx0, **tin)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(); s.solve(tx0, **tin); b.record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev])
    print("   kernel ms p50 %.4f min %.4f" % (np.percentile(t, 50), t.min()))
    for Bs in (1, 32):
        s1 = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=Bs, variant=variant, **W.CTRL_TT)
        t1 = {k: torch.as_tensor(w[k][:Bs]).to(dev) for k in keys}; x1 = torch.as_tensor(w["x0"][:Bs]).to(dev)
        for _ in range(5): s1.solve(x1, **t1)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        for a, b in ev:
            a.record(); s1.solve(x1, **t1); b.record()
        torch.cuda.synchronize()
        print("   B=%d kernel ms p50 %.4f" % (Bs, np.percentile([a.elapsed_time(b) for a, b in ev], 50)))
        s1.close()
    s.close()
