"""Intra-SM / inter-SM interference of the controller kernels: B copies of ONE QP (identical iteration counts), kernel
time against the number of busy warps per SM and busy SMs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N = 8
w0 = W.controller_batch(4096, N, seed=0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
dev = torch.device("cuda", 0)
pick = int(os.environ.get("PICK", "2"))
for variant in [int(v) for v in (sys.argv[1:] or ["6", "8"])]:
    per_task = 4 if variant == 6 else 2
    for B in (per_task, 2 * per_task, 16, 16 * 8, 16 * 148, 32 * 148, 4096):
        w = {k: np.repeat(w0[k][pick:pick + 1], B, axis=0) for k in keys + ("x0",)}
        s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT)
        tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"]).to(dev)
        for _ in range(3): r = s.solve(tx0, **tin)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            a.record(); r = s.solve(tx0, **tin); b.record()
        torch.cuda.synchronize()
        t = np.percentile([a.elapsed_time(b) for a, b in ev], 50)
        print("variant %d B %5d iters %d: %.4f ms" % (variant, B, int(r.iters[0]), t))
        s.close()
