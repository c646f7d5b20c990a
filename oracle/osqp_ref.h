/* TEST INFRASTRUCTURE — interface of the CPU OSQP restatement (see osqp_ref.c header).
 * Status integers are OSQP's own. */
#ifndef OSQP_REF_H
#define OSQP_REF_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
  OSQP_REF_SOLVED = 1,
  OSQP_REF_SOLVED_INACCURATE = 2,
  OSQP_REF_PRIMAL_INFEASIBLE_INACCURATE = 3,
  OSQP_REF_DUAL_INFEASIBLE_INACCURATE = 4,
  OSQP_REF_MAX_ITER_REACHED = -2,
  OSQP_REF_PRIMAL_INFEASIBLE = -3,
  OSQP_REF_DUAL_INFEASIBLE = -4,
  OSQP_REF_NON_CVX = -7,
  OSQP_REF_UNSOLVED = -10,
  /* return codes of osqp_ref_solve (not statuses) */
  OSQP_REF_DATA_VALIDATION_ERROR = 1,
  OSQP_REF_LINSYS_ERROR = 4
};

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, delta, adaptive_rho_tolerance;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
  int polish, polish_refine_iter, scaled_termination;
  int linsys;         /* 0: sparse LDL' + minimum degree (QDLDL/AMD stand-in); 1: dense LU, partial pivoting */
  int cache_ordering; /* reuse the fill-reducing permutation when the KKT pattern repeats */
} osqp_ref_settings;

typedef struct {
  /* caller-allocated outputs; optional ones may be NULL */
  double *x;  /* n  unscaled primal (NaN if no solution) */
  double *y;  /* m  unscaled dual */
  double *z;  /* m  unscaled A x image (optional) */
  double *xs, *zs, *ys; /* scaled-space ADMM iterates before polish (optional) */
  unsigned char *active_lo, *active_up; /* m each, polish active-set guess (optional) */
  double *D, *E, *Ps, *As, *qs;         /* scaling + scaled data (optional) */
  int status, iter, rho_updates, status_polish, n_factor;
  double obj_val, pri_res, dua_res, rho_estimate, rho_final, c;
} osqp_ref_result;

void osqp_ref_default_settings(osqp_ref_settings *s);

/* P: upper-triangular csc (n x n); A: csc (m x n).  Returns 0 or an OSQP_REF_*_ERROR code. */
int osqp_ref_solve(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                   const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                   const osqp_ref_settings *settings, osqp_ref_result *res);

#ifdef __cplusplus
}
#endif
#endif
