"""How the state L2 is left in before a launch changes the scheduling kernel's time: (a) 1 GiB memset (dirty lines to write
back), (b) 1 GiB read (clean lines), (c) nothing (the previous launch's own output), each with the GPU kept busy in front
of the launch so that the event pair brackets the kernel only."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lpvmpc_b200 as lp
W = lp.workloads
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 8
dev = torch.device("cuda", 0)
track = lp.Map("L_shape")
w = W.controller_batch(B, N, seed=0, track=track)
solver = lp.BatchSolver("controller", N, W.CTRL_DT, track=track.PointAndTangent, max_batch=B, device=0, **W.CTRL_TT)
keys = ("u_prev", "vel_ref", "curv_ref", "lap")
tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}
tin["x0"] = torch.as_tensor(w["x0"]).to(dev)
flush = torch.empty(1024 * 1024 * 1024, dtype=torch.uint8, device=dev)
flush.zero_()
for _ in range(5): solver.schedule(**tin)
torch.cuda.synchronize()
def run(mode, steps=20):
    st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        if mode == "memset": flush.zero_()
        elif mode == "read": flush.view(torch.int64).sum()
        elif mode == "memset+read": flush.zero_(); flush.view(torch.int64)[: 1 << 25].sum()
        else: torch.cuda._sleep(400000)
        st[i].record(); solver.schedule(**tin); en[i].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in zip(st, en)])
    return float(np.percentile(t, 50)), float(t.mean())
for mode in ("memset", "read", "memset+read", "none", "memset"):
    p50, mean = run(mode)
    print(json.dumps({"B": B, "l2_state": mode, "kernel_us_p50": round(p50 * 1e3, 1), "kernel_us_mean": round(mean * 1e3, 1), "GBps_p50": round(3776 * B / (p50 * 1e-3) * 1e-9)}))
