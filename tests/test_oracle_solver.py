"""Guard rails for the OSQP restatement (oracle/osqp_ref.c).  Upstream `osqp` is an absent, un-pinned PyPI
dependency of the reference ("parity unpinned"), so the restatement is anchored by
solver-independent facts:

  * the worked example of the OSQP documentation / paper (known optimum),
  * KKT optimality of every converged solution in the exact (unscaled) problem data,
  * an active-set re-solve of the same QP with a dense LU,
  * analytically infeasible / unbounded problems must come back with the right certificates,
  * the two linear-system back-ends of the oracle (no-pivot sparse LDL' as QDLDL does, pivoted dense LU) agree.
"""
import os

import numpy as np
import pytest
from scipy import sparse

import oracle
import tests_common as tc

INF = 1e30


def _golden_qps(golden_dir):
    """(name, P, q, A, l, u) of every QP the reference's own Python handed to osqp.setup in make_golden.py."""
    out = []
    for fname, prefix in (("controller.npz", "c"), ("planner.npz", "p")):
        g = np.load(os.path.join(golden_dir, fname), allow_pickle=False)
        for i in range(int(g["n_cases"])):
            for qp in ("qp_", "wu_qp_"):
                p = "%s%d_%s" % (prefix, i, qp)
                if p + "shape" not in g:
                    continue
                m, n = g[p + "shape"]
                P = sparse.csc_matrix((g[p + "P_data"], g[p + "P_indices"], g[p + "P_indptr"]), shape=(n, n))
                A = sparse.csc_matrix((g[p + "A_data"], g[p + "A_indices"], g[p + "A_indptr"]), shape=(m, n))
                out.append((p, P, g[p + "q"], A, np.maximum(g[p + "l"], -INF), np.minimum(g[p + "u"], INF)))
    return out


def _full_sym(P):
    P = sparse.csc_matrix(P)
    U = sparse.triu(P, format="csc")
    return (U + sparse.triu(P, 1, format="csc").T).toarray()


def _kkt_violation(P, q, A, l, u, x, y):
    """(stationarity, primal feasibility, complementarity) inf-norm residuals in the unscaled data."""
    Pd, Ad = _full_sym(P), A.toarray()
    stat = np.abs(Pd @ x + q + Ad.T @ y).max()
    z = Ad @ x
    prim = max(np.maximum(z - u, 0).max(), np.maximum(l - z, 0).max())
    yp, ym = np.maximum(y, 0), np.minimum(y, 0)
    fin_u, fin_l = u < INF * 1e-4, l > -INF * 1e-4
    comp = max(np.abs(yp[fin_u] * (u - z)[fin_u]).max(initial=0.0), np.abs(ym[fin_l] * (z - l)[fin_l]).max(initial=0.0),
               np.abs(yp[~fin_u]).max(initial=0.0), np.abs(ym[~fin_l]).max(initial=0.0))
    return stat, prim, comp


def test_osqp_documentation_example():
    # OSQP docs "Setup and solve" example / paper demo: optimum x = (0.3, 0.7), y = (-2.9, 0, 0.2), obj = 1.88
    P = sparse.csc_matrix([[4.0, 1.0], [1.0, 2.0]])
    q = np.array([1.0, 1.0])
    A = sparse.csc_matrix([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])
    l, u = np.array([1.0, 0.0, 0.0]), np.array([1.0, 0.7, 0.7])
    r = oracle.osqp_solve(P, q, A, l, u, polish=1)
    assert r["status"] == 1 and r["status_polish"] == 1
    np.testing.assert_allclose(r["x"], [0.3, 0.7], atol=1e-9)
    np.testing.assert_allclose(r["y"], [-2.9, 0.0, 0.2], atol=1e-7)
    assert abs(r["obj_val"] - 1.88) < 1e-9
    assert r["iter"] % 25 == 0 and r["iter"] <= 200   # terminates on a check_termination boundary
    # without polish the answer is only eps-accurate, as upstream
    r0 = oracle.osqp_solve(P, q, A, l, u, polish=0)
    assert r0["status"] == 1 and np.abs(r0["x"] - [0.3, 0.7]).max() < 5e-3


def test_golden_qps_kkt_and_active_set_resolve(golden_dir):
    qps = _golden_qps(golden_dir)
    assert len(qps) >= 20
    n_solved = 0
    for name, P, q, A, l, u in qps:
        r = oracle.osqp_solve(P, q, A, l, u, polish=1)
        assert r["status"] in (1, 2, -2, -3, 3), (name, r["status"])
        if r["status"] != 1:
            continue
        n_solved += 1
        scale = max(1.0, np.abs(q).max())
        stat, prim, comp = _kkt_violation(P, q, A, l, u, r["x"], r["y"])
        if r["status_polish"] == 1:
            assert stat < 1e-7 * scale and prim < 1e-8 and comp < 1e-6 * scale, (name, stat, prim, comp)
            # active-set re-solve: with the polish active set the reduced KKT system gives the same point
            act = (r["active_lo"] | r["active_up"]).astype(bool)
            Ad, Pd = A.toarray()[act], _full_sym(P)
            b = np.where(r["active_lo"][act].astype(bool), l[act], u[act])
            n, ma = Pd.shape[0], Ad.shape[0]
            K = np.block([[Pd + 1e-9 * np.eye(n), Ad.T], [Ad, -1e-9 * np.eye(ma)]])
            sol = np.linalg.solve(K, np.concatenate([-q, b]))
            for _ in range(3):
                res = np.concatenate([-q, b]) - np.block([[Pd, Ad.T], [Ad, np.zeros((ma, ma))]]) @ sol
                sol = sol + np.linalg.solve(K, res)
            np.testing.assert_allclose(sol[:n], r["x"], atol=1e-6 * max(1.0, np.abs(r["x"]).max()), err_msg=name)
        else:
            # unpolished: eps-accurate only (eps_abs = eps_rel = 1e-3)
            Ad, Pd = A.toarray(), _full_sym(P)
            eps_p = 1e-3 + 1e-3 * np.abs(Ad @ r["x"]).max()
            eps_d = 1e-3 + 1e-3 * max(np.abs(Pd @ r["x"]).max(), np.abs(Ad.T @ r["y"]).max(), np.abs(q).max())
            assert prim < 1.001 * eps_p and stat < 1.001 * eps_d, (name, stat, prim)
    assert n_solved >= len(qps) // 2


def test_backends_agree_fixed_iterations(golden_dir):
    """No-pivot LDL' (what QDLDL does) vs pivoted dense LU: same ADMM iterates to 1e-9 relative after 100 fixed
    iterations — this is the accuracy yardstick for the condensed block factorisation on the GPU as well."""
    for name, P, q, A, l, u in _golden_qps(golden_dir)[::3]:
        kw = dict(max_iter=100, check_termination=0, adaptive_rho=0, polish=0)
        a = oracle.osqp_solve(P, q, A, l, u, linsys=0, **kw)
        b = oracle.osqp_solve(P, q, A, l, u, linsys=1, **kw)
        for k in ("xs", "zs", "ys"):
            den = max(np.abs(b[k]).max(), 1e-300)
            assert np.abs(a[k] - b[k]).max() / den < 1e-9, (name, k)


def test_degenerate_dual_labels(golden_dir):
    """Rows whose dual is zero up to round-off (dynamics rows of the unweighted state s, Q[4,4] = 0) get their
    lower/upper active label from the sign of noise: the two back-ends may disagree there and only there."""
    for name, P, q, A, l, u in _golden_qps(golden_dir)[:8]:
        a = oracle.osqp_solve(P, q, A, l, u, linsys=0, polish=1)
        b = oracle.osqp_solve(P, q, A, l, u, linsys=1, polish=1)
        assert a["status"] == b["status"] and a["iter"] == b["iter"]
        if a["status"] != 1:
            continue
        ys = np.abs(a["ys"])
        firm = ys > 1e-9 * max(ys.max(), 1e-300)
        assert np.array_equal(a["active_lo"][firm], b["active_lo"][firm]), name
        assert np.array_equal(a["active_up"][firm], b["active_up"][firm]), name


def test_primal_infeasible_certificate():
    # x <= -1 and x >= 1
    P = sparse.csc_matrix(np.eye(2))
    A = sparse.csc_matrix([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    l, u = np.array([-INF, 1.0, -1.0]), np.array([-1.0, INF, 1.0])
    r = oracle.osqp_solve(P, np.zeros(2), A, l, u, polish=1)
    assert r["status"] == -3 and np.isnan(r["x"]).all() and r["obj_val"] >= INF


def test_dual_infeasible_certificate():
    # min -x1 s.t. x1 >= 0 (unbounded), P = 0 on x1
    P = sparse.csc_matrix(np.diag([0.0, 1.0]))
    A = sparse.csc_matrix(np.eye(2))
    l, u = np.array([0.0, -1.0]), np.array([INF, 1.0])
    r = oracle.osqp_solve(P, np.array([-1.0, 0.0]), A, l, u, polish=1)
    assert r["status"] == -4 and r["obj_val"] <= -INF


def test_planner_infeasible_initial_speed(golden_dir):
    """Planner x0 is both pinned by the dynamics equality and box-bounded (LPV_MPC_Planner.py:176-181):
    vx0 below min_vel makes the QP primal infeasible (SURVEY.md 8c guard rail 3)."""
    track = np.load(os.path.join(golden_dir, "track.npz"))["L_shape_PointAndTangent"]
    g = np.load(os.path.join(golden_dir, "planner.npz"), allow_pickle=False)
    N = int(g["p0_N"])
    cfg = oracle.make_cfg("planner", N, 1.0 / 20.0, tc.PLAN_Q, tc.PLAN_R, tc.PLAN_DR, track, L_cf=tc.PLAN_L)
    x0 = np.array(g["p0_x"], dtype=np.float64).copy()
    x0[0] = 0.5  # < min_vel = 0.9
    st = oracle.default_settings(polish=1)
    o = oracle.plan_solve(cfg, st, x0, A=g["p0_A"], B=g["p0_B"], mode=0, max_ey=float(g["p0_max_ey"]))
    assert o["status"] in (-3, 3)
    assert np.isnan(o["xPred"]).all()


def test_max_iter_and_inaccurate_statuses():
    rng = np.random.default_rng(0)
    n, m = 12, 18
    M = rng.standard_normal((n, n))
    P = sparse.csc_matrix(M @ M.T + 1e-3 * np.eye(n))
    A = sparse.csc_matrix(rng.standard_normal((m, n)))
    l, u = -rng.uniform(0.1, 1, m), rng.uniform(0.1, 1, m)
    q = rng.standard_normal(n)
    r = oracle.osqp_solve(P, q, A, l, u, max_iter=3, polish=0)
    assert r["status"] in (-2, 2) and r["iter"] == 3
    r = oracle.osqp_solve(P, q, A, l, u, polish=1)
    assert r["status"] == 1
    stat, prim, comp = _kkt_violation(P, q, A, l, u, r["x"], r["y"])
    assert prim < 1e-3 and stat < 1e-2


def test_adaptive_rho_interval_is_deterministic(golden_dir):
    """adaptive_rho_interval = 0 -> 4 * check_termination (deterministic fallback, not wall-clock)."""
    name, P, q, A, l, u = _golden_qps(golden_dir)[0]
    a = oracle.osqp_solve(P, q, A, l, u, polish=0)
    b = oracle.osqp_solve(P, q, A, l, u, polish=0, adaptive_rho_interval=100)
    assert a["iter"] == b["iter"] and a["rho_updates"] == b["rho_updates"]
    np.testing.assert_array_equal(a["x"], b["x"])
