#!/usr/bin/env python
"""bench.py — LPV-MPC QP solves/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (LPV scheduling -> QP build -> OSQP ADMM solve + polish) over one batch of
synthetic problems.  Workloads:
    ctrl4096   4,096 controller QPs, N=8, trajectory-tracking tune, lap=1       (BASELINE configs[1], default)
    plan16384  16,384 planner QPs, N=40, lateral-box "obstacles"                 (configs[2])
    ctrl1024N100  1,024 controller QPs, N=100                                     (configs[4])
    ctrl512N160   512 controller QPs, N=160: the factor exceeds shared memory and is streamed by TMA (not a BASELINE config)
    mc8192     Monte-Carlo closed loop, 8,192 vehicles per GPU (configs[3] is 65,536 over 8 GPUs): a step is
               `--ticks-per-step` controller ticks of the whole fleet, each tick = simulate + localise + schedule +
               build + solve on the device (lpvmpc_loop_*); 24 steps x 23 ticks = the 552-tick lap
Multi-GPU: one process per GPU (torchrun), every rank solves its own batch of the same size (weak scaling, no
collective on the data path); value = all ranks' QPs / max-over-ranks device time.

`value`   : kernel-path throughput, inputs resident in HBM, CUDA-event timed per step (L2 flushed between steps).
`e2e`     : same metric through the public host API (numpy in -> numpy out): pinned staging + H2D + kernel + D2H.
`roofline`: fp64-FMA roofline of the one kernel in the step (algorithmic flops from measured iteration counts).
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference loop (C build + OSQP restatement; the
reference's own OSQP is an absent PyPI dependency), OpenMP over the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic flop model per QP (1 FMA = 2 flops; condensed block-tridiagonal form)
FLOP_TABLE = {
    # (kind, N): F_scale, F_form, F_fac, F_solve, F_iter, F_check
    ("controller", 8): (14e3, 1.6e3, 7.3e3, 2.6e3, 5.6e3, 2.5e3),
    ("controller", 20): (34e3, 4.0e3, 18.3e3, 6.4e3, 13.8e3, 6.0e3),
    ("controller", 100): (168e3, 19.8e3, 91.3e3, 31.7e3, 68.2e3, 29.7e3),
    ("planner", 40): (62e3, 7.1e3, 22.2e3, 9.2e3, 23.1e3, 11.1e3),
}
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r1_fp64_peak.jsonl")
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # {workload: {"variant": v, "dram_bytes_per_launch": b, "source": ...}}
KERNEL_NAMES = {1: "lpv_solve_kernel (generic warp-per-QP)", 2: "lpv_solve_t8_kernel", 3: "lpv_solve_g8_kernel",
                5: "lpv_solve_h8_kernel (factor in shared memory)", 6: "lpv_solve_h8t_kernel (factor in tensor memory)",
                7: "lpv_solve_h8_kernel (factor streamed from the L2 slab by TMA bulk copies, ring of 4 stage blocks)",
                8: "lpv_solve_h16t_kernel (16 lanes per QP, twisted block factor in tensor memory)"}

WORKLOADS = {
    "ctrl4096": dict(kind="controller", N=8, B=4096, seed=0),
    "plan16384": dict(kind="planner", N=40, B=16384, seed=1, harvest=True),
    # round 1's planner batch (nominal roll-outs, half width 0.3): kept for comparison with the round-1 numbers
    "plan16384nominal": dict(kind="planner", N=40, B=16384, seed=1),
    "ctrl1024N100": dict(kind="controller", N=100, B=1024, seed=3),
    # not a BASELINE config: the cfg-2 distribution at a batch that fills every QP slot of the GPU many times over
    # (what a Monte-Carlo tick of configs[3] looks like to the solver: 8,192 vehicles per GPU and more)
    "ctrl65536": dict(kind="controller", N=8, B=65536, seed=0),
    "mc8192": dict(kind="fleet", N=8, B=8192, seed=2),
    # not a BASELINE config: a horizon whose block factor (165 KB) no longer fits shared memory next to the stage vectors,
    # so the factor is streamed from the L2 slab by TMA bulk copies (kernel variant 7)
    "ctrl512N160": dict(kind="controller", N=160, B=512, seed=3),
    # not a BASELINE config: the planner main loop (plannerMain.py:128-224) for a fleet of plans, i.e. planner QPs with the
    # reference's own closed-loop distribution (SURVEY 8d cfg 3 generator); a step = `--ticks-per-step` re-planning ticks
    "planloop4096": dict(kind="planfleet", N=40, B=4096, seed=5),
    # not a BASELINE config: the stand-alone LPVPrediction kernel (lpvmpc_schedule_*: A_k, B_k and the roll-out written to
    # HBM) on the cfg-2 distribution — the HBM-bound kernel of SURVEY 8d; 65,536 QPs = 250 MB per launch (> L2)
    "sched65536": dict(kind="schedule", N=8, B=65536, seed=0),
}


OSQP_NOTE = "defaults (eps 1e-3, rho 0.1 adaptive@100, sigma 1e-6, alpha 1.6, 10 Ruiz passes, max_iter 4000) + polish"


def base_config(name, ticks_per_step=None):
    """The workload description both arms print (identical keys and values: the driver compares them)."""
    spec = WORKLOADS[name]
    c = {"workload": name, "kind": spec["kind"], "N": spec["N"], "batch_per_gpu": spec["B"], "seed": spec["seed"], "osqp": OSQP_NOTE}
    if spec["kind"] in ("fleet", "planfleet"):
        c["ticks_per_step"] = ticks_per_step
    return c


def flops_per_qp(kind, N, iters, rho_updates, polished):
    key = (kind, N)
    if key not in FLOP_TABLE:  # scale the nearest row per stage
        base = min((k for k in FLOP_TABLE if k[0] == kind), key=lambda k: abs(k[1] - N))
        f = [v * N / base[1] for v in FLOP_TABLE[base]]
    else:
        f = FLOP_TABLE[key]
    F_scale, F_form, F_fac, F_solve, F_iter, F_check = f
    n_fac = 1.0 + rho_updates
    w = F_scale + F_form + n_fac * F_fac + iters * F_iter + np.ceil(iters / 25.0) * F_check
    w = w + polished * (F_fac + 4 * F_solve)
    return w


def fp64_peak_tflops():
    try:
        with open(FP64_PEAK_FILE) as fh:
            for line in fh:
                d = json.loads(line)
                if d.get("bench") == "dfma_sustained":
                    return float(d["tflops"]), "profiles/r1_fp64_peak.jsonl (own DFMA microbenchmark, sustained 3 s)"
    except Exception:
        pass
    return 34.2, "fallback: DFMA microbenchmark of round 1"


def ncu_traffic(workload, variant):
    """DRAM bytes per launch of the solve kernel from the committed ncu --set full capture (None if there is none)."""
    try:
        with open(NCU_TRAFFIC_FILE) as fh:
            d = json.load(fh).get(workload)
        if d and int(d.get("variant", -1)) == int(variant):
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {"hbm_gbs": 6650.0, "_fallback": True}


def init_nccl_quietly(dist, dev):
    """init_process_group with file descriptor 1 pointed at stderr: NCCL writes its version banner to stdout when the
    communicator is created, and stdout must carry exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        import torch
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler(object):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        return {"sm_mhz": (float(np.median(self.samples)) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_workload(name, rank):
    import lpvmpc_b200 as lp
    W = lp.workloads
    spec = WORKLOADS[name]
    track = lp.Map("L_shape")
    seed = spec["seed"] + 1000 * rank
    if spec["kind"] == "controller":
        w = W.controller_batch(spec["B"], spec["N"], seed=seed, track=track, steer_scale=(0.2 if spec["N"] >= 50 else 1.0))
        tune, dt = W.CTRL_TT, W.CTRL_DT
        keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
    else:
        # BASELINE configs[2] as SURVEY 8d words it: tuples harvested from the reference's own planner loop, perturbed
        w = W.planner_batch_harvest(spec["B"], spec["N"], seed=seed) if spec.get("harvest") else W.planner_batch(spec["B"], spec["N"], seed=seed, track=track)
        tune, dt = W.PLAN, W.PLAN_DT
        keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
    return spec, track, w, tune, dt, keys


def cpu_reference_rate(name, threads, repeats=1, sample=None):
    """QP/s of the CPU oracle port on `threads` host threads, on (a sample of) the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    spec, track, w, tune, dt, keys = make_workload(name, 0)
    B = spec["B"] if sample is None else min(sample, spec["B"])
    st = oracle.default_settings(polish=1)
    if spec["kind"] == "controller":
        cfg = oracle.make_cfg("controller", spec["N"], dt, tune["Q"], tune["R"], tune["dR"], track.PointAndTangent)
        run = lambda: oracle.ctrl_batch(cfg, st, w["x0"][:B], w["u_prev"][:B], w["vel_ref"][:B], w["curv_ref"][:B],
                                        w["lap"][:B], w["u_old"][:B], threads=threads)
    else:
        cfg = oracle.make_cfg("planner", spec["N"], dt, tune["Q"], tune["R"], tune["dR"], track.PointAndTangent, L_cf=tune["L_cf"])
        run = lambda: oracle.plan_batch(cfg, st, w["x0"][:B], w["SS"][:B], w["u_prev"][:B], w["u_old"][:B], w["max_ey"][:B],
                                        w["ey_lo"][:B], w["ey_hi"][:B], threads=threads)
    times = []
    solved = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = run()
        times.append(time.perf_counter() - t0)
        solved = int(r["solved"])
    return B, times, solved


def cpu_fleet_rate(threads, vehicles, ticks, seed=2):
    """vehicle-ticks/s of the CPU oracle's closed loop (oracle/loop_ref.c) on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    import lpvmpc_b200 as lp
    W = lp.workloads
    m = lp.Map("L_shape")
    cfg = oracle.make_cfg("controller", 8, W.CTRL_DT, W.CTRL_PT["Q"], W.CTRL_PT["R"], W.CTRL_PT["dR"], m.PointAndTangent)
    st = oracle.default_settings(polish=1)
    lc = oracle.loop_cfg(half_width=m.halfWidth, slack=m.slack)
    state = oracle.loop_state(lp.fleet_start(vehicles, seed=seed, track_map=m), 8)
    t0 = time.perf_counter()
    solved = oracle.loop_run(cfg, st, lc, state, ticks, threads=threads)
    dt = time.perf_counter() - t0
    done = int(state["ctr"][:, 7].sum())
    return done / dt, dt, solved / float(max(done, 1))


def run_fleet_reference(args):
    threads = os.cpu_count() or 1
    vehicles, ticks = max(64, 8 * threads), args.ticks_per_step
    for _ in range(max(1, args.warmup)):   # full-size untimed passes: the host's OpenMP threads take a while to spin up
        cpu_fleet_rate(threads, vehicles, ticks)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, sf = cpu_fleet_rate(threads, vehicles, ticks)
        rates.append(r); times.append(dt)
    value = vehicles * ticks * len(times) / float(np.sum(times))
    line = {
        "impl": "reference", "metric": "LPV-MPC QP solves/sec", "value": value, "unit": "QP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.workload, ticks),
        "note": "CPU oracle port of the closed loop (controller main loop + Simulator.f + getLocalPosition + OSQP restatement) on a bounded "
                "sample fleet of %d vehicles, every step restarts it from tick 0; the reference's Python overhead is NOT included" % vehicles,
        "cpu_baseline": {"value": value, "unit": "QP/s", "cores": threads, "kind": "port",
                         "sample": "%d vehicles x %d ticks per step x %d steps" % (vehicles, ticks, len(times))},
        "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solved_fraction": sf,
    }
    print(json.dumps(line))
    return 0


def run_fleet(args):
    """configs[3]: the device-resident closed loop.  A step = args.ticks_per_step ticks of the whole fleet."""
    import torch
    import lpvmpc_b200 as lp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        init_nccl_quietly(dist, dev)
    spec = WORKLOADS[args.workload]
    B, tps = spec["B"], args.ticks_per_step
    m = lp.Map("L_shape")
    sim0 = lp.fleet_start(B, seed=spec["seed"] + 1000 * rank, track_map=m)
    fleet = lp.ClosedLoopFleet(m, N=spec["N"], max_fleet=B, device=local, variant=args.variant)
    info0 = fleet.solver.info()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream(local).cuda_stream
    t0 = torch.as_tensor(sim0).to(dev)
    fleet.start(t0)
    for _ in range(args.warmup):   # warm-up steps also carry the fleet past the 9 _EstimateABC ticks
        fleet.run(tps, stream=stream)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    launches0 = fleet.solver.info()["kernel_launches"]
    before = fleet.read(("stat", "ctr"))
    sampler.start()
    barrier()
    for i in range(args.steps):
        starts[i].record()
        fleet.run(tps, stream=stream)
        ends[i].record()
    barrier()
    clocks = sampler.stop()
    launches = fleet.solver.info()["kernel_launches"] - launches0
    after = fleet.read(("stat", "ctr"))
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = float(step_ms.sum())
    ticks_done = float((after["ctr"][:, 7] - before["ctr"][:, 7]).sum())
    solved = float((after["stat"][:, 0] - before["stat"][:, 0]).sum())
    iters_sum = float((after["stat"][:, 1] - before["stat"][:, 1]).sum())
    # algorithmic flops: every tick is one controller QP with the measured iteration count, 1 factorisation (+ polish)
    F_scale, F_form, F_fac, F_solve, F_iter, F_check = FLOP_TABLE[("controller", spec["N"])]
    flops = ticks_done * (F_scale + F_form + F_fac + (F_fac + 4 * F_solve)) + iters_sum * (F_iter + F_check / 25.0)

    # end to end through the host API: numpy start states in, whole run, numpy fleet state out
    barrier()
    te = time.perf_counter()
    fleet.start(sim0)
    fleet.run(tps * args.steps)
    out = fleet.read()
    e2e_total = time.perf_counter() - te
    barrier()
    e2e_ticks = float(out["ctr"][:, 7].sum())
    h2d = int(sim0.nbytes)
    d2h = int(sum(v.nbytes for v in out.values()))

    if dist is not None:
        t = torch.tensor([total_ms, e2e_total, ticks_done, solved, e2e_ticks], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_total = float(tmax[0]), float(tmax[1])
        ticks_all, solved_all, e2e_ticks_all = float(tsum[2]), float(tsum[3]), float(tsum[4])
    else:
        ticks_all, solved_all, e2e_ticks_all = ticks_done, solved, e2e_ticks
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak, peak_src = fp64_peak_tflops()
    ms_per_step = total_ms / args.steps
    achieved = flops / (total_ms * 1e-3) * 1e-12
    line = {
        "metric": "LPV-MPC QP solves/sec", "value": ticks_all / (total_ms * 1e-3), "unit": "QP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.workload, tps),
        "setup": {"what": "closed loop: Simulator.f x7 + getLocalPosition + LPVPrediction + build + OSQP + polish per tick",
                  "tune": "path tracking (controllerMain.py:139-141)", "l2": "fleet state + solver slab stay resident by design (closed loop); no flush",
                  "kernel_variant": info0["variant"], "swap_ey_epsi": 1},
        "e2e": {"value": e2e_ticks_all / e2e_total, "unit": "QP/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                "ms_per_step": 1e3 * e2e_total / args.steps, "note": "start(host states) + run(steps x ticks) + read(): one H2D and one D2H per run, nothing per tick"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": KERNEL_NAMES.get(info0["variant"], "?") + " (one launch per tick) + lpv_loop_kernel",
                     "algorithmic_flops_per_step": flops / args.steps},
        "kernel_latency_ms": {"per_tick_p50": float(np.percentile(step_ms, 50)) / tps, "per_tick_p99": float(np.percentile(step_ms, 99)) / tps},
        "solved_fraction": solved_all / max(ticks_all, 1.0),
        "iters": {"mean": iters_sum / max(ticks_done, 1.0)},
        "retired_vehicles": int((after["ctr"][:, 5] != 0).sum()),
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        vehicles = max(64, 8 * threads)
        cpu_fleet_rate(threads, vehicles, 4 * tps)
        r, dt, _ = cpu_fleet_rate(threads, vehicles, 4 * tps)
        line["cpu_baseline"] = {"value": r, "unit": "QP/s", "cores": threads, "kind": "port",
                                "sample": "%d vehicles x %d ticks from the same start distribution, OpenMP over all host cores (%.1f s)" % (vehicles, 4 * tps, dt)}
    print(json.dumps(line))
    fleet.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def planfleet_start(B, seed):
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(1.0, 2.0, B), rng.normal(0, 0.01, B), rng.normal(0, 0.05, B), rng.normal(0, 0.02, B), rng.normal(0, 0.02, B)], axis=1)
    return x0, rng.uniform(0.0, 18.0, B)


def run_planfleet(args):
    """Planner main loop for a fleet (lpvmpc_plan_loop_*): single GPU, device-resident; a step = ticks_per_step ticks."""
    import torch
    import lpvmpc_b200 as lp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("planloop4096 is a single-GPU workload")
    spec = WORKLOADS[args.workload]
    B, N = spec["B"], spec["N"]
    tps = min(args.ticks_per_step, 4)
    m = lp.Map("L_shape")
    x0, s0 = planfleet_start(B, spec["seed"])
    fleet = lp.PlannerFleet(m, N=N, max_fleet=B, max_ey=0.2, variant=args.variant)
    info0 = fleet.solver.info()
    stream = torch.cuda.current_stream(0).cuda_stream
    fleet.start(x0, s0)
    for _ in range(args.warmup):
        fleet.run(tps, stream=stream)
    torch.cuda.synchronize()
    before = fleet.read(("stat", "ctr"))
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = fleet.solver.info()["kernel_launches"]
    e0.record()
    for _ in range(args.steps):
        fleet.run(tps, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    after = fleet.read(("stat", "ctr"))
    ticks_done = float((after["ctr"][:, 0] - before["ctr"][:, 0]).sum())
    solved = float((after["stat"][:, 0] - before["stat"][:, 0]).sum())
    iters_sum = float((after["stat"][:, 1] - before["stat"][:, 1]).sum())
    F_scale, F_form, F_fac, F_solve, F_iter, F_check = FLOP_TABLE[("planner", N)]
    flops = ticks_done * (F_scale + F_form + 1.5 * F_fac) + iters_sum * (F_iter + F_check / 25.0)
    peak, peak_src = fp64_peak_tflops()
    te = time.perf_counter()
    fleet.start(x0, s0)
    fleet.run(tps * min(args.steps, 3))
    out = fleet.read()
    e2e_total = time.perf_counter() - te
    line = {
        "metric": "LPV-MPC QP solves/sec", "value": ticks_done / (total_ms * 1e-3), "unit": "QP/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "kind": "planner fleet (plannerMain loop: re-plan from xPred[1], arc-length integration)", "N": N,
                   "plans_per_gpu": B, "ticks_per_step": tps, "kernel_variant": info0["variant"], "l2": "fleet state resident by design; no flush"},
        "e2e": {"value": float(out["ctr"][:, 0].sum()) / e2e_total, "unit": "QP/s", "h2d_bytes_per_step": int(x0.nbytes + s0.nbytes) // min(args.steps, 3),
                "d2h_bytes_per_step": int(sum(v.nbytes for v in out.values())) // min(args.steps, 3),
                "note": "start(host states) + run + read(): from tick 0 (includes the slow first ticks)"},
        "gpu_launches": int(fleet.solver.info()["kernel_launches"] - launches0), "clocks": clocks,
        "roofline": {"bound": "fp64_fma", "achieved": flops / (total_ms * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                     "frac": flops / (total_ms * 1e-3) * 1e-12 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": KERNEL_NAMES.get(info0["variant"], "?") + " (one launch per tick) + lpv_plan_loop_kernel"},
        "solved_fraction": solved / max(ticks_done, 1.0), "iters": {"mean": iters_sum / max(ticks_done, 1.0)},
        "retired_plans": int((after["ctr"][:, 3] != 0).sum()),
    }
    print(json.dumps(line))
    fleet.close()
    return 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if WORKLOADS[args.workload]["kind"] == "planfleet":
        print(json.dumps({"impl": "reference", "unavailable": "planloop4096 is not a BASELINE config; use plan16384"}))
        return 0
    if WORKLOADS[args.workload]["kind"] == "schedule":
        print(json.dumps({"impl": "reference", "unavailable": "sched65536 is not a BASELINE config (scheduling only); its cpu_baseline is on the ours line"}))
        return 0
    if WORKLOADS[args.workload]["kind"] == "fleet":
        return run_fleet_reference(args)
    threads = os.cpu_count() or 1
    spec = WORKLOADS[args.workload]
    sample = spec["B"] if spec["kind"] == "controller" and spec["N"] <= 20 else max(64, threads * 8)
    for _ in range(max(args.warmup, 1) if sample <= 4096 else 1):
        cpu_reference_rate(args.workload, threads, 1, sample=min(sample, 256))
    B, times, solved = cpu_reference_rate(args.workload, threads, args.steps, sample=sample)
    total = float(np.sum(times))
    value = B * len(times) / total
    line = {
        "impl": "reference", "metric": "LPV-MPC QP solves/sec", "value": value, "unit": "QP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.workload),
        "note": "CPU oracle port (C restatement of the reference's LPVPrediction + QP build + OSQP 0.6 algorithm) on %d QPs of the batch per step; "
                "upstream osqp is an absent PyPI dependency; Python overhead of the reference (about 1.7 ms/QP at N=8) is NOT included, "
                "which favours this arm" % B,
        "cpu_baseline": {"value": value, "unit": "QP/s", "cores": threads, "kind": "port",
                         "sample": "%d QPs per step x %d steps (%s of the workload batch)" % (B, len(times), "all" if B == spec["B"] else "a slice")},
        "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solved_fraction": solved / float(B),
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import lpvmpc_b200 as lp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        init_nccl_quietly(dist, dev)

    spec, track, w, tune, dt, keys = make_workload(args.workload, rank)
    B = spec["B"]
    solver = lp.BatchSolver(spec["kind"], spec["N"], dt, track=track.PointAndTangent, max_batch=B, device=local, variant=args.variant, **tune)
    info0 = solver.info()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident arm
    tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}
    tx0 = torch.as_tensor(w["x0"]).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for _ in range(args.warmup):
        r = solver.solve(tx0, **tin)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    launches0 = solver.info()["kernel_launches"]
    sampler.start()
    barrier()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        r = solver.solve(tx0, **tin)
        ends[i].record()
    barrier()
    clocks = sampler.stop()
    launches = solver.info()["kernel_launches"] - launches0
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = float(step_ms.sum())
    status = r.status.cpu().numpy()
    iters = r.iters.cpu().numpy().astype(np.float64)
    rho_up = r.rho_updates.cpu().numpy().astype(np.float64)
    pol = (r.polish_status.cpu().numpy() != 0).astype(np.float64)
    solved = int((status == 1).sum())
    flops = float(flops_per_qp(spec["kind"], spec["N"], iters, rho_up, pol).sum())

    # ---------------------------------------------------------------- end-to-end arm (host API)
    # host arrays in -> host arrays out.  The result arrays are views of the handle's pinned result arenas (two, used in turn:
    # lpvmpc_solve_host_view), which removes the last host copy of the call; `e2e_fresh_arrays` times the same call with the
    # results copied into freshly allocated numpy arrays (lpvmpc_solve_host)
    hin = {k: w[k] for k in keys}
    for _ in range(max(1, min(args.warmup, 3))):
        solver.solve(w["x0"], **hin)
        solver.solve(w["x0"], host_views=True, **hin)
    barrier()
    fresh_t = []
    for i in range(max(3, args.steps // 3)):
        t0 = time.perf_counter()
        rh = solver.solve(w["x0"], **hin)
        fresh_t.append(time.perf_counter() - t0)
    barrier()
    e2e_t = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        rh = solver.solve(w["x0"], host_views=True, **hin)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_total = float(np.sum(e2e_t))
    h2d = int(w["x0"].nbytes + sum(np.asarray(w[k]).nbytes for k in keys))
    d2h = int(sum(rh[k].nbytes for k in rh if hasattr(rh[k], "nbytes")))
    assert np.array_equal(rh.status, status), "host and device paths disagree"

    # ---------------------------------------------------------------- the other BASELINE configs (all ranks take part)
    configs = None
    if args.workload == "ctrl4096" and not args.no_configs:
        gloo = None
        if dist is not None:
            try:
                gloo = dist.new_group(backend="gloo")   # host-side gather of the sharded results ("final host gather")
            except Exception:
                gloo = None
        ctx = dict(rank=rank, world=world, local=local, dev=dev, dist=dist, barrier=barrier, flush=flush, gloo=gloo)
        configs = secondary_configs(args, ctx)

    # ---------------------------------------------------------------- reduce over ranks (max time)
    if dist is not None:
        t = torch.tensor([total_ms, e2e_total, float(solved), float(flops)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_total = float(tmax[0]), float(tmax[1])
        solved_all, flops_rank = float(tsum[2]), flops
    else:
        solved_all = float(solved)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B * args.steps / e2e_total
    peak, peak_src = fp64_peak_tflops()
    achieved = flops / (ms_per_step * 1e-3) * 1e-12  # this rank's kernel: flops per launch / launch duration
    in_bytes_qp = h2d / float(B)
    out_bytes_qp = d2h / float(B)
    mp = measured_peaks()
    line = {
        "metric": "LPV-MPC QP solves/sec", "value": value, "unit": "QP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.workload),
        "setup": {"sched": "fused LPVPrediction", "l2": "flushed between steps (256 MiB memset outside the per-step event pairs)",
                  "kernel_variant": info0["variant"], "workspace_in_smem": info0["workspace_in_smem"],
                  "smem_bytes_per_qp": info0["smem_bytes_per_qp"]},
        "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_total / args.steps,
                "h2d": ("kernel reads its inputs from the pinned host arena the call packed them into (zero-copy: the bytes cross PCIe inside the launch)"
                        if (os.environ.get("LPVMPC_ZERO_COPY_IN", "1" if info0["variant"] == 8 else "0") != "0") else "one cudaMemcpyAsync in front of the kernel"),
                "d2h": ("kernel writes the results into the pinned host arena (zero-copy, posted PCIe writes behind the compute)"
                        if os.environ.get("LPVMPC_ZERO_COPY_OUT", "1") != "0" else "one cudaMemcpyAsync after the kernel"),
                "results": "numpy views of the handle's pinned result arenas (lpvmpc_solve_host_view; two arenas used in turn)",
                "latency_ms": {"p50": 1e3 * float(np.percentile(e2e_t, 50)), "p99": 1e3 * float(np.percentile(e2e_t, 99)),
                               "max": 1e3 * float(np.max(e2e_t))},
                "fresh_arrays_ms_per_step_rank0": 1e3 * float(np.mean(fresh_t))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload, info0["variant"]), "peak_source": peak_src,
                     "kernel": KERNEL_NAMES.get(info0["variant"], "?") + ": schedule + build + Ruiz + factor + ADMM + polish, one launch per step",
                     "algorithmic_flops_per_launch": flops,
                     "algorithmic_hbm_bytes_per_launch": (in_bytes_qp + out_bytes_qp) * B,
                     "hbm_frac_of_measured": ((in_bytes_qp + out_bytes_qp) * B / (ms_per_step * 1e-3) * 1e-9) / float(mp.get("hbm_gbs", 6650.0))},
        "kernel_latency_ms": {"p50": float(np.percentile(step_ms, 50)), "p99": float(np.percentile(step_ms, 99)), "max": float(step_ms.max())},
        "solved_fraction": solved_all / float(world * B),
        "iters": {"mean": float(iters.mean()), "p50": float(np.percentile(iters, 50)), "p99": float(np.percentile(iters, 99)), "max": float(iters.max())},
    }
    if configs is not None:
        line["configs"] = configs
    # the same kernel with every QP slot of the GPU filled many times over (ctrl4096 fills them 1.7 times: its second
    # round is 73 % full); reported beside the headline, not instead of it
    if world == 1 and args.workload == "ctrl4096" and not args.no_saturated:
        spec2, track2, w2, tune2, dt2, keys2 = make_workload("ctrl65536", rank)
        s2 = lp.BatchSolver(spec2["kind"], spec2["N"], dt2, track=track2.PointAndTangent, max_batch=spec2["B"], device=local, variant=args.variant, **tune2)
        tin2 = {k: torch.as_tensor(w2[k]).to(dev) for k in keys2}
        tx2 = torch.as_tensor(w2["x0"]).to(dev)
        for _ in range(3):
            r2 = s2.solve(tx2, **tin2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 5
        e0.record()
        for _ in range(nrep):
            r2 = s2.solve(tx2, **tin2)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / nrep
        it2 = r2.iters.cpu().numpy().astype(np.float64)
        fl2 = float(flops_per_qp(spec2["kind"], spec2["N"], it2, r2.rho_updates.cpu().numpy().astype(np.float64),
                                 (r2.polish_status.cpu().numpy() != 0).astype(np.float64)).sum())
        line["saturated"] = {"workload": "ctrl65536 (same distribution, 65,536 QPs resident in HBM; inputs 21 MB + outputs 39 MB > L2 share, no flush)",
                             "value": spec2["B"] / (ms2 * 1e-3), "unit": "QP/s", "ms_per_step": ms2,
                             "roofline_frac": fl2 / (ms2 * 1e-3) * 1e-12 / peak,
                             "solved_fraction": float((r2.status.cpu().numpy() == 1).mean())}
        s2.close()
    # CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload on all host cores
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = B if spec["kind"] == "controller" and spec["N"] <= 20 else max(64, threads * 8)
        cpu_reference_rate(args.workload, threads, 1, sample=min(sample, 256))
        Bs, times, _ = cpu_reference_rate(args.workload, threads, 3 if sample <= 4096 else 1, sample=sample)
        line["cpu_baseline"] = {"value": Bs * len(times) / float(np.sum(times)), "unit": "QP/s", "cores": threads, "kind": "port",
                                "sample": "%d QPs x %d passes of the same workload, OpenMP over all host cores; C port without the reference's Python overhead" % (Bs, len(times))}
        # the same port on ONE core (the reference loop is single-threaded), and the reference's real Python loop as
        # measured in the build container (tools/ref_python_baseline.py; /root/reference does not exist on this box)
        s1 = min(Bs, 1024 if spec["N"] <= 20 else 32)
        B1, t1, _ = cpu_reference_rate(args.workload, 1, 1, sample=s1)
        line["cpu_baseline_1core"] = {"value": B1 / float(np.sum(t1)), "unit": "QP/s", "cores": 1, "kind": "port", "sample": "%d QPs of the same workload, one thread" % B1}
        try:
            with open(os.path.join(ROOT, "profiles", "r3_cpu_reference_python.json")) as fh:
                rp = json.load(fh)
            line["cpu_reference_python"] = {"value": rp["qp_per_s_1core"], "unit": "QP/s", "cores": 1, "kind": "reference Python build + oracle OSQP",
                                            "ms_per_qp": rp["ms_per_qp"], "measured": "build container, not this box (profiles/r3_cpu_reference_python.json)",
                                            "applies_to": "ctrl4096"} if args.workload == "ctrl4096" else None
        except Exception:
            pass
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_schedule(args):
    """Stand-alone scheduling kernel (LPVPrediction for a batch, matrices materialised): HBM roofline."""
    import torch
    import lpvmpc_b200 as lp
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("sched65536 is a single-GPU workload")
    W = lp.workloads
    spec = WORKLOADS[args.workload]
    B, N = spec["B"], spec["N"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    track = lp.Map("L_shape")
    w = W.controller_batch(B, N, seed=spec["seed"], track=track)
    solver = lp.BatchSolver("controller", N, W.CTRL_DT, track=track.PointAndTangent, max_batch=B, device=0, **W.CTRL_TT)
    keys = ("u_prev", "vel_ref", "curv_ref", "lap")
    tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}
    tin["x0"] = torch.as_tensor(w["x0"]).to(dev)
    # 1 GiB flush (~170 us of memset): besides emptying L2 it keeps the GPU busy while the host prepares the next call, so the
    # event pair brackets the kernel and not the ~40 us of Python / ctypes argument packing in front of a 75 us launch
    flush = torch.empty(1024 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        r = solver.schedule(**tin)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(0)
    launches0 = solver.info()["kernel_launches"]
    sampler.start()
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        r = solver.schedule(**tin)
        ends[i].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = solver.info()["kernel_launches"] - launches0
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    ms = float(step_ms.mean())
    # algorithmic bytes per QP: inputs x0 (6) + u_prev (2N) + vel_ref (N+1) + curv_ref (N) doubles + lap (int32);
    # outputs A_k (36N) + B_k (12N) + roll-out states (6N) doubles + sched_err (int32)
    in_b = 8 * (6 + 2 * N + (N + 1) + N) + 4
    out_b = 8 * (36 * N + 12 * N + 6 * N) + 4
    mp = measured_peaks()
    peak = float(mp.get("hbm_gbs", 6650.0))
    achieved = (in_b + out_b) * B / (ms * 1e-3) * 1e-9
    # end to end: numpy in -> numpy out through lpvmpc_schedule_host
    hin = {k: w[k] for k in keys}
    hin["x0"] = w["x0"]
    for _ in range(2):
        solver.schedule(**hin)
    t0 = time.perf_counter()
    ne = max(3, min(args.steps, 10))
    for _ in range(ne):
        rh = solver.schedule(**hin)
    e2e_s = (time.perf_counter() - t0) / ne
    assert np.array_equal(rh["A_out"], r["A_out"].cpu().numpy()), "host and device paths disagree"
    line = {
        "metric": "LPV-MPC QP solves/sec", "value": B / (ms * 1e-3), "unit": "QP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "kind": "stand-alone LPVPrediction (scheduling only: A_k, B_k, roll-out states to HBM; value = QPs SCHEDULED per second, no solve)",
                   "N": N, "batch_per_gpu": B, "l2": "flushed between steps (1 GiB memset outside the per-step event pairs); 250 MB per launch > L2",
                   "kernel": "lpv_schedule_naive_kernel" if os.environ.get("LPVMPC_SCHED_NAIVE", "0") != "0" else "lpv_schedule_tma_kernel (LPVMPC_SCHED_MODE, default 2)"},
        "e2e": {"value": B / e2e_s, "unit": "QP/s", "h2d_bytes_per_step": int(sum(np.asarray(v).nbytes for v in hin.values())),
                "d2h_bytes_per_step": int(sum(v.nbytes for v in rh.values() if hasattr(v, "nbytes"))), "ms_per_step": 1e3 * e2e_s},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("sched65536", 0) if (B == 65536 and os.environ.get("LPVMPC_SCHED_MODE", "2") == "2") else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if not mp.get("_fallback") else "fallback (B200_PROFILING.md)",
                     "kernel": {"2": "lpv_schedule_tma_kernel (TMA tensor stores)", "1": "lpv_schedule_kernel<CONTROLLER, DIRECT> (256-bit stores)",
                                "0": "lpv_schedule_kernel<CONTROLLER> (tile-staged stores)"}.get(os.environ.get("LPVMPC_SCHED_MODE", "2"), "?"),
                     "algorithmic_bytes_per_qp": in_b + out_b,
                     "algorithmic_bytes_per_launch": (in_b + out_b) * B},
        "kernel_latency_ms": {"p50": float(np.percentile(step_ms, 50)), "max": float(step_ms.max())},
    }
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track.PointAndTangent)
        ns = 4096
        t0 = time.perf_counter()
        for b in range(ns):
            oracle.ctrl_predict(cfg, w["x0"][b], w["u_prev"][b], w["vel_ref"][b], w["curv_ref"][b], 60.0, int(w["lap"][b]))
        line["cpu_baseline"] = {"value": ns / (time.perf_counter() - t0), "unit": "QP/s", "cores": 1, "kind": "port",
                                "sample": "%d QPs of the same batch, one C call of the oracle's LPVPrediction per QP from Python (ctypes overhead included)" % ns}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# The other BASELINE configs, measured inside the default run so that they land in the driver's records
# (VERDICT r1 item 4): compact versions of the per-workload runs above.
def _reduce(dist, dev, vals_max, vals_sum):
    """max / sum over the ranks of two float lists (identity without a process group)."""
    import torch
    if dist is None:
        return list(vals_max), list(vals_sum)
    a = torch.tensor(list(vals_max), dtype=torch.float64, device=dev)
    b = torch.tensor(list(vals_sum), dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return a.tolist(), b.tolist()


def cfg_solve(name, ctx, steps, warmup, sharded, variant=0, e2e=True):
    """One solve workload, device-timed (L2 flushed between steps).  sharded: the workload's B problems are split over
    the ranks by contiguous index range (sharding.shard_range), every rank solves its slice, and the end-to-end leg
    goes host arrays -> sharding.solve_sharded -> gather_results on rank 0 (the "final host gather")."""
    import torch
    import lpvmpc_b200 as lp
    rank, world, local, dev, dist, barrier, flush = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist"], ctx["barrier"], ctx["flush"]
    spec, track, w, tune, dt, keys = make_workload(name, 0 if sharded else rank)
    Btot = spec["B"]
    lo, hi = lp.sharding.shard_range(Btot, rank, world) if sharded else (0, Btot)
    B = hi - lo
    solver = lp.BatchSolver(spec["kind"], spec["N"], dt, track=track.PointAndTangent, max_batch=max(B, 1), device=local, variant=variant, **tune)
    info0 = solver.info()
    tin = {k: torch.as_tensor(np.ascontiguousarray(w[k][lo:hi])).to(dev) for k in keys}
    tx0 = torch.as_tensor(np.ascontiguousarray(w["x0"][lo:hi])).to(dev)
    for _ in range(warmup):
        r = solver.solve(tx0, **tin)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches0 = solver.info()["kernel_launches"]
    for i in range(steps):
        flush.zero_()
        starts[i].record()
        r = solver.solve(tx0, **tin)
        ends[i].record()
    barrier()
    launches = solver.info()["kernel_launches"] - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(starts, ends)])
    status = r.status.cpu().numpy()
    iters = r.iters.cpu().numpy().astype(np.float64)
    flops = float(flops_per_qp(spec["kind"], spec["N"], iters, r.rho_updates.cpu().numpy().astype(np.float64),
                               (r.polish_status.cpu().numpy() != 0).astype(np.float64)).sum()) if B else 0.0
    e2e_s, gathered_ok = None, None
    if e2e:
        hin = {k: w[k] for k in keys}
        gdist = ctx["gloo"] if world > 1 else None
        for rep in range(2):   # first pass warms the pinned staging path
            barrier()
            t0 = time.perf_counter()
            if sharded:
                full = lp.sharding.solve_sharded(solver, Btot, w["x0"], dist=dist, group=gdist, **hin)
            else:
                full = solver.solve(w["x0"], **hin)
            barrier()
            e2e_s = time.perf_counter() - t0
        if rank == 0:
            st_full = np.asarray(full["status"])
            gathered_ok = bool(st_full.shape[0] == (Btot if sharded else B) and np.array_equal(st_full[lo:hi], status))
    peak, _ = fp64_peak_tflops()
    mx, sm = _reduce(dist, dev, [float(step_ms.sum()), e2e_s or 0.0], [float((status == 1).sum()), float(np.isin(status, (1, 2, -2)).sum()), float(iters.sum()), float(B), flops, float(launches)])
    total_ms = mx[0]
    ms = total_ms / steps
    out = {"workload": name, "kind": spec["kind"], "N": spec["N"], "batch_total": int(sm[3]), "scaling": "strong" if sharded else "weak",
           "sharded": bool(sharded), "value": sm[3] * steps / (total_ms * 1e-3), "unit": "QP/s", "ms_per_step": ms, "steps": steps,
           "kernel_variant": info0["variant"], "gpu_launches": int(sm[5]),
           "roofline": {"bound": "fp64_fma", "achieved": sm[4] / world / (ms * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                        "frac": sm[4] / world / (ms * 1e-3) * 1e-12 / peak, "note": "mean over ranks of flops per launch / max-over-ranks launch time"},
           "solved_fraction": sm[0] / max(sm[3], 1.0), "feasible_fraction": sm[1] / max(sm[3], 1.0), "iters_mean": sm[2] / max(sm[3], 1.0),
           "l2": "flushed between steps"}
    if e2e:
        out["e2e"] = {"value": sm[3] / mx[1], "unit": "QP/s", "ms_per_step": 1e3 * mx[1],
                      "path": ("host arrays -> sharding.solve_sharded (shard_range slice per rank, lpvmpc_solve_host) -> gather_results on rank 0 (gloo group)"
                               if sharded else "host arrays -> lpvmpc_solve_host"), "gathered_matches_device": gathered_ok}
    solver.close()
    return out


def cfg_fleet(ctx, ticks, variant=0):
    """configs[3]: 8,192 vehicles per GPU (65,536 on 8), one step of `ticks` controller ticks after the warm-up ticks."""
    import torch
    import lpvmpc_b200 as lp
    rank, world, local, dev, dist, barrier = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist"], ctx["barrier"]
    spec = WORKLOADS["mc8192"]
    B = spec["B"]
    m = lp.Map("L_shape")
    sim0 = lp.fleet_start(B, seed=spec["seed"] + 1000 * rank, track_map=m)
    fleet = lp.ClosedLoopFleet(m, N=spec["N"], max_fleet=B, device=local, variant=variant)
    stream = torch.cuda.current_stream(local).cuda_stream
    fleet.start(torch.as_tensor(sim0).to(dev))
    fleet.run(ticks, stream=stream)       # carries the fleet past the 9 _EstimateABC ticks
    barrier()
    before = fleet.read(("stat", "ctr"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = fleet.solver.info()["kernel_launches"]
    barrier()
    e0.record()
    fleet.run(ticks, stream=stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = fleet.solver.info()["kernel_launches"] - launches0
    after = fleet.read(("stat", "ctr"))
    ticks_done = float((after["ctr"][:, 7] - before["ctr"][:, 7]).sum())
    solved = float((after["stat"][:, 0] - before["stat"][:, 0]).sum())
    iters_sum = float((after["stat"][:, 1] - before["stat"][:, 1]).sum())
    F_scale, F_form, F_fac, F_solve, F_iter, F_check = FLOP_TABLE[("controller", spec["N"])]
    flops = ticks_done * (F_scale + F_form + F_fac + (F_fac + 4 * F_solve)) + iters_sum * (F_iter + F_check / 25.0)
    peak, _ = fp64_peak_tflops()
    mx, sm = _reduce(dist, dev, [ms], [ticks_done, solved, iters_sum, flops, float(launches), float((after["ctr"][:, 5] != 0).sum())])
    out = {"workload": "mc8192", "kind": "fleet", "N": spec["N"], "vehicles_total": B * world, "ticks_per_step": ticks, "scaling": "weak",
           "value": sm[0] / (mx[0] * 1e-3), "unit": "QP/s (vehicle-ticks/s)", "ms_per_step": mx[0], "ms_per_tick": mx[0] / ticks, "steps": 1,
           "kernel_variant": fleet.solver.info()["variant"], "gpu_launches": int(sm[4]),
           "roofline": {"bound": "fp64_fma", "achieved": sm[3] / world / (mx[0] * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                        "frac": sm[3] / world / (mx[0] * 1e-3) * 1e-12 / peak},
           "solved_fraction": sm[1] / max(sm[0], 1.0), "iters_mean": sm[2] / max(sm[0], 1.0), "retired_vehicles": int(sm[5]),
           "l2": "fleet state resident by design; no flush"}
    fleet.close()
    return out


def cfg_schedule(ctx, steps=10, warmup=3):
    """sched65536 on rank 0's GPU: the stand-alone LPVPrediction kernel against the HBM roofline."""
    import torch
    import lpvmpc_b200 as lp
    local, dev = ctx["local"], ctx["dev"]
    W = lp.workloads
    spec = WORKLOADS["sched65536"]
    B, N = spec["B"], spec["N"]
    track = lp.Map("L_shape")
    w = W.controller_batch(B, N, seed=spec["seed"], track=track)
    solver = lp.BatchSolver("controller", N, W.CTRL_DT, track=track.PointAndTangent, max_batch=B, device=local, **W.CTRL_TT)
    keys = ("u_prev", "vel_ref", "curv_ref", "lap")
    tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys}
    tin["x0"] = torch.as_tensor(w["x0"]).to(dev)
    flush = torch.empty(1024 * 1024 * 1024, dtype=torch.uint8, device=dev)   # see run_schedule
    for _ in range(warmup):
        solver.schedule(**tin)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        flush.zero_()
        starts[i].record()
        solver.schedule(**tin)
        ends[i].record()
    torch.cuda.synchronize()
    del flush
    ms = float(np.mean([a.elapsed_time(b) for a, b in zip(starts, ends)]))
    in_b = 8 * (6 + 2 * N + (N + 1) + N) + 4
    out_b = 8 * (36 * N + 12 * N + 6 * N) + 4
    mp = measured_peaks()
    peak = float(mp.get("hbm_gbs", 6650.0))
    achieved = (in_b + out_b) * B / (ms * 1e-3) * 1e-9
    solver.close()
    return {"workload": "sched65536", "kind": "schedule", "N": N, "batch_total": B, "value": B / (ms * 1e-3), "unit": "QPs scheduled/s", "ms_per_step": ms,
            "steps": steps, "gpu_launches": steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("sched65536", 0), "algorithmic_bytes_per_launch": (in_b + out_b) * B,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if not mp.get("_fallback") else "fallback (B200_PROFILING.md)"},
            "l2": "flushed between steps (1 GiB memset); 250 MB per launch > L2"}


def cfg_latency(ctx, variant=0, reps=50):
    """configs[0]'s use case: ONE controller QP per call (controllerMain.py:329-331).  Kernel latency by CUDA events and
    end to end through the host API, beside one CPU core of the oracle port on the same QPs."""
    import torch
    import lpvmpc_b200 as lp
    local, dev = ctx["local"], ctx["dev"]
    W = lp.workloads
    track = lp.Map("L_shape")
    w = W.controller_batch(64, 8, seed=0, track=track)
    keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
    out = {}
    for B in (1, 32):
        solver = lp.BatchSolver("controller", 8, W.CTRL_DT, track=track.PointAndTangent, max_batch=B, device=local, variant=variant, **W.CTRL_TT)
        tin = {k: torch.as_tensor(w[k][:B]).to(dev) for k in keys}
        tx0 = torch.as_tensor(w["x0"][:B]).to(dev)
        for _ in range(5):
            r = solver.solve(tx0, **tin)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record()
            r = solver.solve(tx0, **tin)
            b.record()
        torch.cuda.synchronize()
        k_ms = np.array([a.elapsed_time(b) for a, b in ev])
        hin = {k: w[k][:B] for k in keys}
        for _ in range(3):
            solver.solve(w["x0"][:B], **hin)
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            solver.solve(w["x0"][:B], **hin)
            t.append(time.perf_counter() - t0)
        out["b%d" % B] = {"kernel_ms_p50": float(np.percentile(k_ms, 50)), "kernel_ms_p99": float(np.percentile(k_ms, 99)),
                          "e2e_ms_p50": 1e3 * float(np.percentile(t, 50)), "e2e_ms_p99": 1e3 * float(np.percentile(t, 99)),
                          "iters": [int(v) for v in r.iters.cpu().numpy()[:4]], "kernel_variant": solver.info()["variant"]}
        solver.close()
    return out


def secondary_configs(args, ctx):
    """Every BASELINE config besides the headline one, inside the default run.  Runs on all ranks (collective)."""
    import torch
    cfgs = {}
    t0 = time.perf_counter()
    # configs[2]: 16,384 planner QPs SHARDED over the ranks (strong scaling) — SURVEY 8d's harvested generator
    cfgs["plan16384"] = cfg_solve("plan16384", ctx, steps=2, warmup=1, sharded=True, variant=args.variant)
    # configs[3]: 8,192 vehicles per GPU x one step of 23 ticks (65,536 vehicles on 8 GPUs)
    cfgs["mc8192"] = cfg_fleet(ctx, args.ticks_per_step, variant=args.variant)
    # configs[4]: N = 100, 1,024 QPs per GPU
    cfgs["ctrl1024N100"] = cfg_solve("ctrl1024N100", ctx, steps=3, warmup=1, sharded=False, variant=args.variant, e2e=False)
    if ctx["rank"] == 0:
        cfgs["sched65536"] = cfg_schedule(ctx)
        cfgs["latency_ctrl_n8"] = cfg_latency(ctx, variant=args.variant)
    ctx["barrier"]()
    cfgs["_seconds"] = time.perf_counter() - t0
    return cfgs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ctrl4096", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-saturated", action="store_true", help="skip the 65,536-QP secondary measurement of the ctrl4096 run")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs) of the default ctrl4096 run")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 = auto)")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch per GPU (profiling runs; not a BASELINE config)")
    ap.add_argument("--ticks-per-step", type=int, default=23, help="mc8192: controller ticks per step (24 x 23 = one lap)")
    args = ap.parse_args()
    if args.batch > 0:
        WORKLOADS[args.workload] = dict(WORKLOADS[args.workload], B=args.batch)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if WORKLOADS[args.workload]["kind"] == "fleet":
        return run_fleet(args)
    if WORKLOADS[args.workload]["kind"] == "planfleet":
        return run_planfleet(args)
    if WORKLOADS[args.workload]["kind"] == "schedule":
        return run_schedule(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
