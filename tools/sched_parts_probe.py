"""Scheduling kernel with some outputs switched off (private probe through the C ABI): where does the time go?"""
import os, sys, json, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lpvmpc_b200 as lp
nat = lp._native
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
L = nat.lib()
for outs in (("A_out", "B_out", "states_out"), ("A_out", "B_out"), ("A_out", "states_out"), ("B_out",), ("A_out",), ()):
    Bq, a, res, keep, _ = solver._prepare(tin, outs, nat.SCHED_PREDICT, False, 1, 60.0)
    err = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(0).cuda_stream
    def call(): nat.check(L.lpvmpc_schedule_dev(solver._h, B, C.byref(a), C.c_void_p(err.data_ptr()), C.c_void_p(stream)), solver._h)
    for _ in range(3): call()
    torch.cuda.synchronize()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(10)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
    for i in range(10):
        flush.zero_(); st[i].record(); call(); en[i].record()
    torch.cuda.synchronize()
    t = np.array([x.elapsed_time(y) for x, y in zip(st, en)])
    print(json.dumps({"B": B, "outputs": list(outs), "kernel_us_p50": round(float(np.percentile(t, 50)) * 1e3, 1)}))
