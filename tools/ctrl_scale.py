"""Kernel time of the controller variants against the batch size (same distribution as ctrl4096)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
dev = torch.device("cuda", 0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
variants = [int(v) for v in (sys.argv[1:] or ["6", "8"])]
for B in (2368, 4096, 4736, 8192, 16384, 65536):
    w = W.controller_batch(B, 8, seed=0)
    tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"]).to(dev)
    for variant in variants:
        s = lp.BatchSolver("controller", 8, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT)
        for _ in range(3): r = s.solve(tx0, **tin)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
        for a, b in ev:
            a.record(); s.solve(tx0, **tin); b.record()
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in ev])
        print(json.dumps({"B": B, "variant": variant, "kernel_ms_p50": round(float(np.percentile(t, 50)), 4), "us_per_qp": round(float(np.percentile(t, 50)) * 1e3 / B, 4),
                          "iters_mean": float(r.iters.double().mean())}))
        s.close()
