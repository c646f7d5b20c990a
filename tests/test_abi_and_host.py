"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/lpvmpc.h declares, the ctypes
structs mirror the header, the library fails loudly without a device, host-side helpers (track table, workloads,
sharding) behave, and the 2-rank gloo path shards + gathers in problem order."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

lp = pytest.importorskip("lpvmpc_b200")
nat = lp._native


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lpvmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpvmpc_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    nat.build()
    L = C.CDLL(nat.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 12
    for sym in declared:
        assert hasattr(L, sym), "liblpvmpc.so does not export %s" % sym
    assert sorted(nat.EXPORTS) == declared
    assert L.lpvmpc_abi_version() == nat.ABI_VERSION


def test_struct_sizes_match_header():
    """sizeof() of the ctypes mirrors vs. the C compiler's view of include/lpvmpc.h."""
    code = ('#include <stdio.h>\n#include "lpvmpc.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lpvmpc_settings), '
            'sizeof(lpvmpc_cfg), sizeof(lpvmpc_info), sizeof(lpvmpc_args), sizeof(lpvmpc_loop_cfg), sizeof(lpvmpc_loop_state), '
            'sizeof(lpvmpc_plan_loop_state));return 0;}\n')
    exe = os.path.join(ROOT, "tests", "_sizes.out")
    try:
        subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code.encode(), check=True)
        out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    finally:
        if os.path.exists(exe):
            os.remove(exe)
    got = [C.sizeof(nat.Settings), C.sizeof(nat.Cfg), C.sizeof(nat.Info), C.sizeof(nat.Args), C.sizeof(nat.LoopCfg),
           C.sizeof(nat.LoopState), C.sizeof(nat.PlanLoopState)]
    assert [int(v) for v in out] == got


def test_default_settings_are_osqp_defaults_plus_polish():
    s = nat.default_settings()
    assert (s.rho, s.sigma, s.alpha) == (0.1, 1e-6, 1.6)
    assert (s.eps_abs, s.eps_rel, s.eps_prim_inf, s.eps_dual_inf) == (1e-3, 1e-3, 1e-4, 1e-4)
    assert (s.max_iter, s.check_termination, s.scaling, s.adaptive_rho, s.polish, s.polish_refine_iter) == (4000, 25, 10, 1, 1, 3)
    with pytest.raises(KeyError):
        nat.default_settings(not_a_setting=1)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    track = lp.Map("L_shape").PointAndTangent
    W = lp.workloads
    with pytest.raises(nat.NativeError) as e:
        lp.BatchSolver("controller", 8, W.CTRL_DT, track=track, max_batch=4, **W.CTRL_TT)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_create_rejects_bad_arguments():
    L = nat.lib()
    h = C.c_void_p()
    assert L.lpvmpc_create(None, C.byref(h)) == -1
    cfg = nat.Cfg()
    cfg.abi_version = 999
    assert L.lpvmpc_create(C.byref(cfg), C.byref(h)) == -1
    assert b"abi_version" in L.lpvmpc_last_error(None)
    cfg.abi_version = nat.ABI_VERSION
    cfg.kind = 7
    assert L.lpvmpc_create(C.byref(cfg), C.byref(h)) == -1
    assert L.lpvmpc_get_info(None, None) == -1
    L.lpvmpc_destroy(None)  # no-op


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "autonomous-racing-lpv-mpp-mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r'#include\s*[<"](osqp_ref|lpv_ref)', txt), f
                assert "liblpv_oracle" not in txt, f


def test_track_table_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "track.npz"))
    m = lp.Map("L_shape")
    np.testing.assert_allclose(m.PointAndTangent, g["L_shape_PointAndTangent"], rtol=0, atol=1e-12)
    assert abs(m.TrackLength - 19.2296) < 1e-3 and abs(m.halfWidth - 0.3) < 1e-12
    for s, kap, ok in zip(g["curv_s"], g["curv_kappa"], g["curv_ok"]):
        if ok:
            assert lp.curvature(float(s), m.PointAndTangent) == kap
        else:
            with pytest.raises(TypeError):
                lp.curvature(float(s), m.PointAndTangent)


def test_workload_generators_are_seeded_and_in_range():
    W = lp.workloads
    a, b = W.controller_batch(64, 8, seed=3), W.controller_batch(64, 8, seed=3)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["x0"].shape == (64, 6) and a["u_prev"].shape == (64, 8, 2) and a["vel_ref"].shape == (64, 9)
    assert np.abs(a["u_prev"][:, :, 0]).max() <= 0.249 and a["vel_ref"].max() <= 5.0
    p = W.planner_batch(32, 40, seed=1)
    assert p["SS"].shape == (32, 41) and (np.diff(p["SS"], axis=1) > 0).all()
    assert (p["ey_lo"] <= p["ey_hi"]).all() and (p["ey_lo"] > -0.3 - 1e-12).all()


def test_shard_ranges_partition_the_batch():
    sh = lp.sharding
    for B in (0, 1, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            r = [sh.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
import lpvmpc_b200 as lp

class FakeSolver(object):
    # host-logic stand-in for BatchSolver (no GPU here): result rows are functions of the input rows only
    def solve(self, x0, **kw):
        return lp.BatchResult(x_pred=np.repeat(x0[:, None, :], 3, axis=1) * 2.0, u_pred=kw["u_prev"] + 1.0,
                              status=np.ones(x0.shape[0], dtype=np.int32), _keepalive=[object()])

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
B = 37
rng = np.random.default_rng(0)
x0 = rng.standard_normal((B, 6)); u_prev = rng.standard_normal((B, 8, 2))
out = lp.sharding.solve_sharded(FakeSolver(), B, x0, dist=dist, u_prev=u_prev)
if dist.get_rank() == 0:
    assert out is not None and out["x_pred"].shape == (B, 3, 6)
    assert np.array_equal(out["x_pred"], np.repeat(x0[:, None, :], 3, axis=1) * 2.0)
    assert np.array_equal(out["u_pred"], u_prev + 1.0) and out["status"].sum() == B
    print("RANK0 OK")
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se.decode()[-2000:]
    assert b"RANK0 OK" in outs[0][0]
