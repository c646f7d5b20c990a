/*
 * TEST INFRASTRUCTURE — CPU restatement of the OSQP ADMM QP solver ("osqp_ref").
 *
 * PARITY UNPINNED: the reference (euge2838/Autonomous-Racing-LPV-MPP-MPC) calls the third-party
 * PyPI package `osqp` (version un-pinned; 0.6.x is the newest that installs on the reference's
 * Python 2.7) at PathFollowingLPVMPC.py:302-323 and LPV_MPC_Planner.py:204-215 with
 * `verbose=False, polish=True` and library defaults otherwise.  That package is neither vendored
 * under /root/reference nor installable in this image, and the reference holds no golden vectors
 * for the solve, so this file restates OSQP 0.6's *published algorithm* (Stellato et al., "OSQP:
 * an operator splitting solver for quadratic programs", Math. Prog. Comp. 2020, Alg. 1 + §5
 * (termination, infeasibility), §5.2 (adaptive rho), §4 (polish), §5.1 (Ruiz equilibration)) in
 * the order of operations of the 0.6 C sources as documented in SURVEY.md Appendix A.5.
 * It is anchored by solver-independent checks in tests/ (KKT optimality, active-set re-solve,
 * analytically infeasible cases), not by upstream golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file; the product path never does.
 *
 * Deliberate, documented deviation from "OSQP as shipped": with adaptive_rho_interval = 0 the
 * shipped library derives the interval from wall-clock time (non-deterministic).  This
 * restatement uses the library's own deterministic fallback (profiling off):
 * interval = ADAPTIVE_RHO_MULTIPLE_TERMINATION(4) * check_termination.
 *
 * Linear system: quasi-definite KKT [[P+sigma I, A'],[A, -diag(1/rho)]] factorised LDL' without
 * pivoting (up-looking sparse LDL', elimination tree, fill-reducing minimum-degree permutation),
 * as QDLDL+AMD do upstream.  linsys=1 swaps in dense LU with partial pivoting as an accuracy
 * yardstick for the tests.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: plain IEEE double ops).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "osqp_ref.h"

#define OSQP_INFTY 1e30
#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define ADAPTIVE_RHO_MULTIPLE_TERMINATION 4
#define ADAPTIVE_RHO_FIXED 100

#define c_max(a, b) (((a) > (b)) ? (a) : (b))
#define c_min(a, b) (((a) < (b)) ? (a) : (b))
#define c_absval(x) (((x) < 0) ? -(x) : (x))

/* ------------------------------------------------------------------ csc helpers */
typedef struct {
  int m, n;
  int *p, *i;
  double *x;
} csc;

static csc *csc_alloc(int m, int n, int nnz) {
  csc *M = (csc *)calloc(1, sizeof(csc));
  M->m = m; M->n = n;
  M->p = (int *)calloc((size_t)n + 1, sizeof(int));
  M->i = (int *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(int));
  M->x = (double *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(double));
  return M;
}
static void csc_free(csc *M) {
  if (!M) return;
  free(M->p); free(M->i); free(M->x); free(M);
}
static csc *csc_copy(int m, int n, const int *p, const int *i, const double *x) {
  csc *M = csc_alloc(m, n, p[n]);
  memcpy(M->p, p, sizeof(int) * ((size_t)n + 1));
  memcpy(M->i, i, sizeof(int) * (size_t)p[n]);
  memcpy(M->x, x, sizeof(double) * (size_t)p[n]);
  return M;
}

/* y (=,+=,-=) A x : plus_eq 0 / 1 / -1, column sweep */
static void mat_vec(const csc *A, const double *x, double *y, int plus_eq) {
  int i, j;
  if (!plus_eq) for (i = 0; i < A->m; i++) y[i] = 0;
  if (A->p[A->n] == 0) return;
  if (plus_eq == -1) {
    for (j = 0; j < A->n; j++)
      for (i = A->p[j]; i < A->p[j + 1]; i++) y[A->i[i]] -= A->x[i] * x[j];
  } else {
    for (j = 0; j < A->n; j++)
      for (i = A->p[j]; i < A->p[j + 1]; i++) y[A->i[i]] += A->x[i] * x[j];
  }
}
/* y (=,+=,-=) A' x, optionally skipping the diagonal (for the strictly-lower part of sym P) */
static void mat_tpose_vec(const csc *A, const double *x, double *y, int plus_eq, int skip_diag) {
  int i, j, k;
  if (!plus_eq) for (i = 0; i < A->n; i++) y[i] = 0;
  if (A->p[A->n] == 0) return;
  if (plus_eq == -1) {
    for (j = 0; j < A->n; j++)
      for (k = A->p[j]; k < A->p[j + 1]; k++) {
        i = A->i[k];
        if (skip_diag && i == j) continue;
        y[j] -= A->x[k] * x[i];
      }
  } else {
    for (j = 0; j < A->n; j++)
      for (k = A->p[j]; k < A->p[j + 1]; k++) {
        i = A->i[k];
        if (skip_diag && i == j) continue;
        y[j] += A->x[k] * x[i];
      }
  }
}
static double vec_norm_inf(const double *v, int l) {
  double mx = 0.0, a;
  for (int i = 0; i < l; i++) { a = c_absval(v[i]); if (a > mx) mx = a; }
  return mx;
}
static double vec_scaled_norm_inf(const double *S, const double *v, int l) {
  double mx = 0.0, a;
  for (int i = 0; i < l; i++) { a = c_absval(S[i] * v[i]); if (a > mx) mx = a; }
  return mx;
}
static double vec_prod(const double *a, const double *b, int n) {
  double p = 0.0;
  for (int i = 0; i < n; i++) p += a[i] * b[i];
  return p;
}
static double quad_form(const csc *P, const double *x) {
  double q = 0.;
  for (int j = 0; j < P->n; j++)
    for (int ptr = P->p[j]; ptr < P->p[j + 1]; ptr++) {
      int i = P->i[ptr];
      if (i == j) q += .5 * P->x[ptr] * x[i] * x[i];
      else if (i < j) q += P->x[ptr] * x[i] * x[j];
    }
  return q;
}
static void mat_inf_norm_cols(const csc *M, double *E) {
  for (int j = 0; j < M->n; j++) E[j] = 0.;
  for (int j = 0; j < M->n; j++)
    for (int p = M->p[j]; p < M->p[j + 1]; p++) E[j] = c_max(c_absval(M->x[p]), E[j]);
}
static void mat_inf_norm_rows(const csc *M, double *E) {
  for (int j = 0; j < M->m; j++) E[j] = 0.;
  for (int j = 0; j < M->n; j++)
    for (int p = M->p[j]; p < M->p[j + 1]; p++) {
      int i = M->i[p];
      E[i] = c_max(c_absval(M->x[p]), E[i]);
    }
}
static void mat_inf_norm_cols_sym_triu(const csc *M, double *E) {
  for (int j = 0; j < M->n; j++) E[j] = 0.;
  for (int j = 0; j < M->n; j++)
    for (int p = M->p[j]; p < M->p[j + 1]; p++) {
      int i = M->i[p];
      double a = c_absval(M->x[p]);
      E[j] = c_max(a, E[j]);
      if (i != j) E[i] = c_max(a, E[i]);
    }
}
static void limit_scaling(double *D, int n) {
  for (int i = 0; i < n; i++) {
    D[i] = D[i] < MIN_SCALING ? 1.0 : D[i];
    D[i] = D[i] > MAX_SCALING ? MAX_SCALING : D[i];
  }
}

/* ------------------------------------------------------------------ linear system back-ends */
typedef struct {
  int dim, n, m;
  int kind;             /* 0 sparse LDL', 1 dense LU */
  /* assembly: upper-triangular permuted KKT in csc + source codes */
  int *Kp, *Ki; double *Kx;
  int *Ksrc;            /* >=0: index into src value arrays, see assemble */
  int knz;
  int *perm, *iperm;    /* perm[new] = old */
  /* LDL' */
  int *Lp, *Li, *Parent, *Lnz, *Flag, *Pattern;
  double *Lx, *D, *Dinv, *Y, *work;
  /* dense LU */
  double *LU; int *piv;
} linsys;

/* Greedy exact minimum-degree ordering on the KKT graph (stands in for AMD). */
static void min_degree_order(int dim, int nnz, const int *ri, const int *ci, int *perm) {
  unsigned char *adj = (unsigned char *)calloc((size_t)dim * dim, 1);
  int *deg = (int *)calloc((size_t)dim, sizeof(int));
  unsigned char *gone = (unsigned char *)calloc((size_t)dim, 1);
  int *nb = (int *)malloc(sizeof(int) * (size_t)dim);
  for (int k = 0; k < nnz; k++) {
    int a = ri[k], b = ci[k];
    if (a != b && !adj[(size_t)a * dim + b]) {
      adj[(size_t)a * dim + b] = adj[(size_t)b * dim + a] = 1;
      deg[a]++; deg[b]++;
    }
  }
  for (int step = 0; step < dim; step++) {
    int best = -1, bd = dim + 1;
    for (int v = 0; v < dim; v++) if (!gone[v] && deg[v] < bd) { bd = deg[v]; best = v; }
    perm[step] = best; gone[best] = 1;
    int cnt = 0;
    unsigned char *row = adj + (size_t)best * dim;
    for (int v = 0; v < dim; v++) if (row[v] && !gone[v]) nb[cnt++] = v;
    for (int a = 0; a < cnt; a++) {
      int va = nb[a];
      adj[(size_t)va * dim + best] = 0; deg[va]--;
      for (int b = a + 1; b < cnt; b++) {
        int vb = nb[b];
        if (!adj[(size_t)va * dim + vb]) {
          adj[(size_t)va * dim + vb] = adj[(size_t)vb * dim + va] = 1;
          deg[va]++; deg[vb]++;
        }
      }
    }
  }
  free(adj); free(deg); free(gone); free(nb);
}

/* per-thread cache of the last ordering (pattern-keyed): the MPC KKT pattern repeats every call */
static __thread int cache_dim = 0, cache_nnz = 0;
static __thread int *cache_ri = 0, *cache_ci = 0, *cache_perm = 0;

static void get_order(int dim, int nnz, const int *ri, const int *ci, int *perm, int use_cache) {
  if (use_cache && cache_dim == dim && cache_nnz == nnz &&
      !memcmp(cache_ri, ri, sizeof(int) * (size_t)nnz) && !memcmp(cache_ci, ci, sizeof(int) * (size_t)nnz)) {
    memcpy(perm, cache_perm, sizeof(int) * (size_t)dim);
    return;
  }
  min_degree_order(dim, nnz, ri, ci, perm);
  if (use_cache) {
    free(cache_ri); free(cache_ci); free(cache_perm);
    cache_ri = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    cache_ci = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    cache_perm = (int *)malloc(sizeof(int) * (size_t)dim);
    memcpy(cache_ri, ri, sizeof(int) * (size_t)nnz); memcpy(cache_ci, ci, sizeof(int) * (size_t)nnz);
    memcpy(cache_perm, perm, sizeof(int) * (size_t)dim);
    cache_dim = dim; cache_nnz = nnz;
  }
}

static void linsys_free(linsys *s) {
  if (!s) return;
  free(s->Kp); free(s->Ki); free(s->Kx); free(s->Ksrc); free(s->perm); free(s->iperm);
  free(s->Lp); free(s->Li); free(s->Parent); free(s->Lnz); free(s->Flag); free(s->Pattern);
  free(s->Lx); free(s->D); free(s->Dinv); free(s->Y); free(s->work); free(s->LU); free(s->piv);
  free(s);
}

/* source codes: [0,pnz) P entry; [pnz, pnz+anz) A entry; then m "minus param2" diagonals;
 * then n "sigma-only" diagonals (P has no stored diagonal there). */
static double src_value(int code, const csc *P, const csc *A, double sigma, const double *param2) {
  int pnz = P->p[P->n], anz = A->p[A->n];
  if (code < pnz) {
    /* diagonal of P gets +sigma: detect by (row==col) at assembly, flagged with code+big */
    return P->x[code];
  }
  if (code < pnz + anz) return A->x[code - pnz];
  if (code < pnz + anz + A->m) return -param2[code - pnz - anz];
  return sigma;
}

static int ldl_numeric(linsys *s);
static int lu_numeric(linsys *s);

/* (re)load numerical values of the KKT and factorise */
static int linsys_refactor(linsys *s, const csc *P, const csc *A, double sigma, const double *param2) {
  int pnz = P->p[P->n];
  for (int k = 0; k < s->knz; k++) {
    int code = s->Ksrc[k];
    if (code >= 0) s->Kx[k] = src_value(code, P, A, sigma, param2);
    else { /* P diagonal entry: value + sigma */
      int c = -code - 1;
      s->Kx[k] = P->x[c] + sigma;
      (void)pnz;
    }
  }
  return s->kind == 0 ? ldl_numeric(s) : lu_numeric(s);
}

static linsys *linsys_init(const csc *P, const csc *A, double sigma, const double *param2, int kind,
                           int cache_order) {
  int n = P->n, m = A->m, dim = n + m;
  int pnz = P->p[n], anz = A->p[A->n];
  linsys *s = (linsys *)calloc(1, sizeof(linsys));
  s->dim = dim; s->n = n; s->m = m; s->kind = kind;
  int maxnz = pnz + anz + m + n;
  int *ri = (int *)malloc(sizeof(int) * (size_t)maxnz), *ci = (int *)malloc(sizeof(int) * (size_t)maxnz);
  int *code = (int *)malloc(sizeof(int) * (size_t)maxnz);
  unsigned char *hasdiag = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
  int nz = 0;
  for (int j = 0; j < n; j++)
    for (int p = P->p[j]; p < P->p[j + 1]; p++) {
      int i = P->i[p];
      ri[nz] = i; ci[nz] = j;
      if (i == j) { code[nz] = -p - 1; hasdiag[j] = 1; } else code[nz] = p;
      nz++;
    }
  for (int j = 0; j < n; j++) if (!hasdiag[j]) { ri[nz] = j; ci[nz] = j; code[nz] = pnz + anz + m + j; nz++; }
  for (int j = 0; j < n; j++)
    for (int p = A->p[j]; p < A->p[j + 1]; p++) { ri[nz] = j; ci[nz] = n + A->i[p]; code[nz] = pnz + p; nz++; }
  for (int i = 0; i < m; i++) { ri[nz] = n + i; ci[nz] = n + i; code[nz] = pnz + anz + i; nz++; }
  free(hasdiag);

  s->perm = (int *)malloc(sizeof(int) * (size_t)dim);
  s->iperm = (int *)malloc(sizeof(int) * (size_t)dim);
  if (kind == 0) get_order(dim, nz, ri, ci, s->perm, cache_order);
  else for (int k = 0; k < dim; k++) s->perm[k] = k;
  for (int k = 0; k < dim; k++) s->iperm[s->perm[k]] = k;

  /* permuted upper-triangular csc */
  s->knz = nz;
  s->Kp = (int *)calloc((size_t)dim + 1, sizeof(int));
  s->Ki = (int *)malloc(sizeof(int) * (size_t)nz);
  s->Kx = (double *)malloc(sizeof(double) * (size_t)nz);
  s->Ksrc = (int *)malloc(sizeof(int) * (size_t)nz);
  for (int k = 0; k < nz; k++) {
    int a = s->iperm[ri[k]], b = s->iperm[ci[k]];
    int col = a > b ? a : b;
    s->Kp[col + 1]++;
  }
  for (int k = 0; k < dim; k++) s->Kp[k + 1] += s->Kp[k];
  int *next = (int *)malloc(sizeof(int) * (size_t)dim);
  memcpy(next, s->Kp, sizeof(int) * (size_t)dim);
  for (int k = 0; k < nz; k++) {
    int a = s->iperm[ri[k]], b = s->iperm[ci[k]];
    int col = a > b ? a : b, row = a > b ? b : a;
    int q = next[col]++;
    s->Ki[q] = row; s->Ksrc[q] = code[k];
  }
  free(next); free(ri); free(ci); free(code);

  if (kind == 0) {
    /* symbolic: elimination tree + column counts of L */
    s->Parent = (int *)malloc(sizeof(int) * (size_t)dim);
    s->Lnz = (int *)malloc(sizeof(int) * (size_t)dim);
    s->Flag = (int *)malloc(sizeof(int) * (size_t)dim);
    s->Pattern = (int *)malloc(sizeof(int) * (size_t)dim);
    s->Lp = (int *)malloc(sizeof(int) * ((size_t)dim + 1));
    for (int k = 0; k < dim; k++) {
      s->Parent[k] = -1; s->Flag[k] = k; s->Lnz[k] = 0;
      for (int p = s->Kp[k]; p < s->Kp[k + 1]; p++) {
        int i = s->Ki[p];
        if (i < k)
          for (; s->Flag[i] != k; i = s->Parent[i]) {
            if (s->Parent[i] == -1) s->Parent[i] = k;
            s->Lnz[i]++;
            s->Flag[i] = k;
          }
      }
    }
    s->Lp[0] = 0;
    for (int k = 0; k < dim; k++) s->Lp[k + 1] = s->Lp[k] + s->Lnz[k];
    int lnz = s->Lp[dim];
    s->Li = (int *)malloc(sizeof(int) * (size_t)(lnz > 0 ? lnz : 1));
    s->Lx = (double *)malloc(sizeof(double) * (size_t)(lnz > 0 ? lnz : 1));
    s->D = (double *)malloc(sizeof(double) * (size_t)dim);
    s->Dinv = (double *)malloc(sizeof(double) * (size_t)dim);
    s->Y = (double *)malloc(sizeof(double) * (size_t)dim);
    s->work = (double *)malloc(sizeof(double) * (size_t)dim);
  } else {
    s->LU = (double *)malloc(sizeof(double) * (size_t)dim * dim);
    s->piv = (int *)malloc(sizeof(int) * (size_t)dim);
    s->work = (double *)malloc(sizeof(double) * (size_t)dim);
  }
  if (linsys_refactor(s, P, A, sigma, param2)) { linsys_free(s); return 0; }
  return s;
}

/* up-looking sparse LDL' (row k of L from a sparse triangular solve over the etree reach) */
static int ldl_numeric(linsys *s) {
  int dim = s->dim;
  for (int k = 0; k < dim; k++) {
    int top = dim;
    s->Y[k] = 0.0; s->Flag[k] = k; s->Lnz[k] = 0;
    for (int p = s->Kp[k]; p < s->Kp[k + 1]; p++) {
      int i = s->Ki[p];
      if (i <= k) {
        s->Y[i] += s->Kx[p];
        int len = 0;
        for (; s->Flag[i] != k; i = s->Parent[i]) { s->Pattern[len++] = i; s->Flag[i] = k; }
        while (len > 0) s->Pattern[--top] = s->Pattern[--len];
      }
    }
    s->D[k] = s->Y[k]; s->Y[k] = 0.0;
    for (; top < dim; top++) {
      int i = s->Pattern[top];
      double yi = s->Y[i];
      s->Y[i] = 0.0;
      int p2 = s->Lp[i] + s->Lnz[i];
      for (int p = s->Lp[i]; p < p2; p++) s->Y[s->Li[p]] -= s->Lx[p] * yi;
      double l_ki = yi * s->Dinv[i];
      s->D[k] -= l_ki * yi;
      s->Li[p2] = k; s->Lx[p2] = l_ki; s->Lnz[i]++;
    }
    if (s->D[k] == 0.0) return -1;
    s->Dinv[k] = 1.0 / s->D[k];
  }
  return 0;
}
static int lu_numeric(linsys *s) {
  int dim = s->dim;
  memset(s->LU, 0, sizeof(double) * (size_t)dim * dim);
  for (int j = 0; j < dim; j++)
    for (int p = s->Kp[j]; p < s->Kp[j + 1]; p++) {
      int i = s->Ki[p];
      s->LU[(size_t)i * dim + j] = s->Kx[p];
      s->LU[(size_t)j * dim + i] = s->Kx[p];
    }
  for (int k = 0; k < dim; k++) {
    int pv = k; double mx = fabs(s->LU[(size_t)k * dim + k]);
    for (int i = k + 1; i < dim; i++) { double a = fabs(s->LU[(size_t)i * dim + k]); if (a > mx) { mx = a; pv = i; } }
    s->piv[k] = pv;
    if (mx == 0.0) return -1;
    if (pv != k) for (int j = 0; j < dim; j++) {
      double t = s->LU[(size_t)k * dim + j]; s->LU[(size_t)k * dim + j] = s->LU[(size_t)pv * dim + j]; s->LU[(size_t)pv * dim + j] = t;
    }
    double inv = 1.0 / s->LU[(size_t)k * dim + k];
    for (int i = k + 1; i < dim; i++) {
      double f = s->LU[(size_t)i * dim + k] * inv;
      if (f != 0.0) {
        s->LU[(size_t)i * dim + k] = f;
        for (int j = k + 1; j < dim; j++) s->LU[(size_t)i * dim + j] -= f * s->LU[(size_t)k * dim + j];
      }
    }
  }
  return 0;
}
/* solve K b' = b in place */
static void linsys_solve_raw(linsys *s, double *b) {
  int dim = s->dim;
  double *w = s->work;
  if (s->kind == 0) {
    for (int k = 0; k < dim; k++) w[k] = b[s->perm[k]];
    for (int j = 0; j < dim; j++) {
      double wj = w[j];
      for (int p = s->Lp[j]; p < s->Lp[j + 1]; p++) w[s->Li[p]] -= s->Lx[p] * wj;
    }
    for (int j = 0; j < dim; j++) w[j] *= s->Dinv[j];
    for (int j = dim - 1; j >= 0; j--) {
      double wj = w[j];
      for (int p = s->Lp[j]; p < s->Lp[j + 1]; p++) wj -= s->Lx[p] * w[s->Li[p]];
      w[j] = wj;
    }
    for (int k = 0; k < dim; k++) b[s->perm[k]] = w[k];
  } else {
    for (int k = 0; k < dim; k++) w[k] = b[k];
    for (int k = 0; k < dim; k++) { /* whole rows (incl. stored multipliers) were swapped: permute first */
      int pv = s->piv[k];
      if (pv != k) { double t = w[k]; w[k] = w[pv]; w[pv] = t; }
    }
    for (int k = 0; k < dim; k++)
      for (int i = k + 1; i < dim; i++) w[i] -= s->LU[(size_t)i * dim + k] * w[k];
    for (int k = dim - 1; k >= 0; k--) {
      double a = w[k];
      for (int j = k + 1; j < dim; j++) a -= s->LU[(size_t)k * dim + j] * w[j];
      w[k] = a / s->LU[(size_t)k * dim + k];
    }
    for (int k = 0; k < dim; k++) b[k] = w[k];
  }
}

/* ------------------------------------------------------------------ solver workspace */
typedef struct {
  int n, m;
  csc *P, *A;
  double *q, *l, *u;
  double *D, *E, *Dinv, *Einv, c, cinv;
  double *rho_vec, *rho_inv_vec; int *constr_type;
  double *x, *y, *z, *xz_tilde, *x_prev, *z_prev;
  double *Ax, *Px, *Aty, *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
  double *D_temp, *D_temp_A, *E_temp;
  double *sol;
  osqp_ref_settings st;
  linsys *ls;
  /* info */
  int iter, status, status_polish, rho_updates, n_factor;
  double obj_val, pri_res, dua_res, rho_estimate;
} work_t;

static double *dvec(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

static void scale_data(work_t *w) {
  int n = w->n, m = w->m;
  w->c = 1.0;
  for (int i = 0; i < n; i++) { w->D[i] = 1.; w->Dinv[i] = 1.; }
  for (int i = 0; i < m; i++) { w->E[i] = 1.; w->Einv[i] = 1.; }
  for (int it = 0; it < w->st.scaling; it++) {
    mat_inf_norm_cols_sym_triu(w->P, w->D_temp);
    mat_inf_norm_cols(w->A, w->D_temp_A);
    for (int i = 0; i < n; i++) w->D_temp[i] = c_max(w->D_temp[i], w->D_temp_A[i]);
    mat_inf_norm_rows(w->A, w->E_temp);
    limit_scaling(w->D_temp, n);
    limit_scaling(w->E_temp, m);
    for (int i = 0; i < n; i++) w->D_temp[i] = sqrt(w->D_temp[i]);
    for (int i = 0; i < m; i++) w->E_temp[i] = sqrt(w->E_temp[i]);
    for (int i = 0; i < n; i++) w->D_temp[i] = 1. / w->D_temp[i];
    for (int i = 0; i < m; i++) w->E_temp[i] = 1. / w->E_temp[i];
    /* P <- D P D : rows first, then columns */
    for (int j = 0; j < n; j++) for (int p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->D_temp[w->P->i[p]];
    for (int j = 0; j < n; j++) for (int p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->D_temp[j];
    /* A <- E A D */
    for (int j = 0; j < n; j++) for (int p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->E_temp[w->A->i[p]];
    for (int j = 0; j < n; j++) for (int p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->D_temp[j];
    for (int i = 0; i < n; i++) w->q[i] = w->D_temp[i] * w->q[i];
    for (int i = 0; i < n; i++) w->D[i] = w->D[i] * w->D_temp[i];
    for (int i = 0; i < m; i++) w->E[i] = w->E[i] * w->E_temp[i];
    /* cost scaling */
    mat_inf_norm_cols_sym_triu(w->P, w->D_temp);
    double c_temp = 0.;
    for (int i = 0; i < n; i++) c_temp += w->D_temp[i];
    c_temp = c_temp / n;
    double inf_norm_q = vec_norm_inf(w->q, n);
    limit_scaling(&inf_norm_q, 1);
    c_temp = c_max(c_temp, inf_norm_q);
    limit_scaling(&c_temp, 1);
    c_temp = 1. / c_temp;
    for (int p = 0; p < w->P->p[n]; p++) w->P->x[p] *= c_temp;
    for (int i = 0; i < n; i++) w->q[i] *= c_temp;
    w->c *= c_temp;
  }
  w->cinv = 1. / w->c;
  for (int i = 0; i < n; i++) w->Dinv[i] = 1. / w->D[i];
  for (int i = 0; i < m; i++) w->Einv[i] = 1. / w->E[i];
  for (int i = 0; i < m; i++) w->l[i] = w->E[i] * w->l[i];
  for (int i = 0; i < m; i++) w->u[i] = w->E[i] * w->u[i];
}

static void set_rho_vec(work_t *w) {
  w->st.rho = c_min(c_max(w->st.rho, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if ((w->l[i] < -OSQP_INFTY * MIN_SCALING) && (w->u[i] > OSQP_INFTY * MIN_SCALING)) {
      w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN;
    } else if (w->u[i] - w->l[i] < RHO_TOL) {
      w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho;
    } else {
      w->constr_type[i] = 0; w->rho_vec[i] = w->st.rho;
    }
    w->rho_inv_vec[i] = 1. / w->rho_vec[i];
  }
}

static void kkt_solve(work_t *w, double *b) {
  /* solve, then x_tilde = sol[:n], z_tilde = b[n:] + rho_inv * nu */
  int n = w->n, m = w->m;
  memcpy(w->sol, b, sizeof(double) * (size_t)(n + m));
  linsys_solve_raw(w->ls, w->sol);
  for (int j = 0; j < n; j++) b[j] = w->sol[j];
  for (int j = 0; j < m; j++) b[j + n] += w->rho_inv_vec[j] * w->sol[j + n];
}

static double compute_pri_res(work_t *w, const double *x, const double *z) {
  mat_vec(w->A, x, w->Ax, 0);
  for (int i = 0; i < w->m; i++) w->z_prev[i] = w->Ax[i] - z[i];
  if (w->st.scaling && !w->st.scaled_termination) return vec_scaled_norm_inf(w->Einv, w->z_prev, w->m);
  return vec_norm_inf(w->z_prev, w->m);
}
static double compute_pri_tol(work_t *w, double eps_abs, double eps_rel) {
  double max_rel_eps, t;
  if (w->st.scaling && !w->st.scaled_termination) {
    max_rel_eps = vec_scaled_norm_inf(w->Einv, w->z, w->m);
    t = vec_scaled_norm_inf(w->Einv, w->Ax, w->m);
    max_rel_eps = c_max(max_rel_eps, t);
  } else {
    max_rel_eps = vec_norm_inf(w->z, w->m);
    t = vec_norm_inf(w->Ax, w->m);
    max_rel_eps = c_max(max_rel_eps, t);
  }
  return eps_abs + eps_rel * max_rel_eps;
}
static double compute_dua_res(work_t *w, const double *x, const double *y) {
  int n = w->n;
  memcpy(w->x_prev, w->q, sizeof(double) * (size_t)n);
  mat_vec(w->P, x, w->Px, 0);
  mat_tpose_vec(w->P, x, w->Px, 1, 1);
  for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Px[i];
  if (w->m > 0) {
    mat_tpose_vec(w->A, y, w->Aty, 0, 0);
    for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Aty[i];
  }
  if (w->st.scaling && !w->st.scaled_termination) return w->cinv * vec_scaled_norm_inf(w->Dinv, w->x_prev, n);
  return vec_norm_inf(w->x_prev, n);
}
static double compute_dua_tol(work_t *w, double eps_abs, double eps_rel) {
  double max_rel_eps, t;
  int n = w->n;
  if (w->st.scaling && !w->st.scaled_termination) {
    max_rel_eps = vec_scaled_norm_inf(w->Dinv, w->q, n);
    t = vec_scaled_norm_inf(w->Dinv, w->Aty, n); max_rel_eps = c_max(max_rel_eps, t);
    t = vec_scaled_norm_inf(w->Dinv, w->Px, n); max_rel_eps = c_max(max_rel_eps, t);
    max_rel_eps *= w->cinv;
  } else {
    max_rel_eps = vec_norm_inf(w->q, n);
    t = vec_norm_inf(w->Aty, n); max_rel_eps = c_max(max_rel_eps, t);
    t = vec_norm_inf(w->Px, n); max_rel_eps = c_max(max_rel_eps, t);
  }
  return eps_abs + eps_rel * max_rel_eps;
}
static double compute_obj_val(work_t *w, const double *x) {
  double o = quad_form(w->P, x) + vec_prod(w->q, x, w->n);
  if (w->st.scaling) o *= w->cinv;
  return o;
}

static int is_primal_infeasible(work_t *w, double eps_prim_inf) {
  int m = w->m;
  double norm_delta_y, ineq_lhs = 0.0;
  for (int i = 0; i < m; i++) {
    if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0;
      else w->delta_y[i] = c_min(w->delta_y[i], 0.0);
    } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) {
      w->delta_y[i] = c_max(w->delta_y[i], 0.0);
    }
  }
  if (w->st.scaling && !w->st.scaled_termination) {
    for (int i = 0; i < m; i++) w->Adelta_x[i] = w->E[i] * w->delta_y[i];
    norm_delta_y = vec_norm_inf(w->Adelta_x, m);
  } else norm_delta_y = vec_norm_inf(w->delta_y, m);
  if (norm_delta_y > eps_prim_inf) {
    for (int i = 0; i < m; i++)
      ineq_lhs += w->u[i] * c_max(w->delta_y[i], 0) + w->l[i] * c_min(w->delta_y[i], 0);
    if (ineq_lhs < -eps_prim_inf * norm_delta_y) {
      mat_tpose_vec(w->A, w->delta_y, w->Atdelta_y, 0, 0);
      if (w->st.scaling && !w->st.scaled_termination)
        for (int i = 0; i < w->n; i++) w->Atdelta_y[i] = w->Dinv[i] * w->Atdelta_y[i];
      return vec_norm_inf(w->Atdelta_y, w->n) < eps_prim_inf * norm_delta_y;
    }
  }
  return 0;
}
static int is_dual_infeasible(work_t *w, double eps_dual_inf) {
  int n = w->n, m = w->m;
  double norm_delta_x, cost_scaling;
  if (w->st.scaling && !w->st.scaled_termination) {
    norm_delta_x = vec_scaled_norm_inf(w->D, w->delta_x, n);
    cost_scaling = w->c;
  } else { norm_delta_x = vec_norm_inf(w->delta_x, n); cost_scaling = 1.0; }
  if (norm_delta_x > eps_dual_inf) {
    if (vec_prod(w->q, w->delta_x, n) < -cost_scaling * eps_dual_inf * norm_delta_x) {
      mat_vec(w->P, w->delta_x, w->Pdelta_x, 0);
      mat_tpose_vec(w->P, w->delta_x, w->Pdelta_x, 1, 1);
      if (w->st.scaling && !w->st.scaled_termination)
        for (int i = 0; i < n; i++) w->Pdelta_x[i] = w->Dinv[i] * w->Pdelta_x[i];
      if (vec_norm_inf(w->Pdelta_x, n) < cost_scaling * eps_dual_inf * norm_delta_x) {
        mat_vec(w->A, w->delta_x, w->Adelta_x, 0);
        if (w->st.scaling && !w->st.scaled_termination)
          for (int i = 0; i < m; i++) w->Adelta_x[i] = w->Einv[i] * w->Adelta_x[i];
        for (int i = 0; i < m; i++) {
          if (((w->u[i] < OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] > eps_dual_inf * norm_delta_x)) ||
              ((w->l[i] > -OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] < -eps_dual_inf * norm_delta_x)))
            return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

static void update_info(work_t *w, int iter) {
  w->iter = iter;
  w->pri_res = w->m == 0 ? 0. : compute_pri_res(w, w->x, w->z);
  w->dua_res = compute_dua_res(w, w->x, w->y);
}

static int check_termination(work_t *w, int approximate) {
  double eps_prim, eps_dual;
  int prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0, dual_inf_check = 0;
  double eps_abs = w->st.eps_abs, eps_rel = w->st.eps_rel;
  double eps_prim_inf = w->st.eps_prim_inf, eps_dual_inf = w->st.eps_dual_inf;
  if ((w->pri_res > OSQP_INFTY) || (w->dua_res > OSQP_INFTY) || isnan(w->pri_res) || isnan(w->dua_res)) {
    /* NB: the NaN test is an addition (a NaN residual can never terminate upstream either; it
       runs to max_iter).  Reported as NON_CVX so that batch callers see a definite status. */
    w->status = OSQP_REF_NON_CVX;
    w->obj_val = NAN;
    return 1;
  }
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_prim_inf *= 10; eps_dual_inf *= 10; }
  if (w->m == 0) prim_res_check = 1;
  else {
    eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (w->pri_res < eps_prim) prim_res_check = 1;
    else prim_inf_check = is_primal_infeasible(w, eps_prim_inf);
  }
  eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
  if (w->dua_res < eps_dual) dual_res_check = 1;
  else dual_inf_check = is_dual_infeasible(w, eps_dual_inf);

  if (prim_res_check && dual_res_check) {
    w->status = approximate ? OSQP_REF_SOLVED_INACCURATE : OSQP_REF_SOLVED;
    return 1;
  } else if (prim_inf_check) {
    w->status = approximate ? OSQP_REF_PRIMAL_INFEASIBLE_INACCURATE : OSQP_REF_PRIMAL_INFEASIBLE;
    if (w->st.scaling && !w->st.scaled_termination)
      for (int i = 0; i < w->m; i++) w->delta_y[i] = w->E[i] * w->delta_y[i];
    w->obj_val = OSQP_INFTY;
    return 1;
  } else if (dual_inf_check) {
    w->status = approximate ? OSQP_REF_DUAL_INFEASIBLE_INACCURATE : OSQP_REF_DUAL_INFEASIBLE;
    if (w->st.scaling && !w->st.scaled_termination)
      for (int i = 0; i < w->n; i++) w->delta_x[i] = w->D[i] * w->delta_x[i];
    w->obj_val = -OSQP_INFTY;
    return 1;
  }
  return 0;
}

static double compute_rho_estimate(work_t *w) {
  int n = w->n, m = w->m;
  double pri_res = vec_norm_inf(w->z_prev, m);
  double dua_res = vec_norm_inf(w->x_prev, n);
  double pri_res_norm = vec_norm_inf(w->z, m);
  double t = vec_norm_inf(w->Ax, m);
  pri_res_norm = c_max(pri_res_norm, t);
  pri_res /= (pri_res_norm + 1e-10);
  double dua_res_norm = vec_norm_inf(w->q, n);
  t = vec_norm_inf(w->Aty, n); dua_res_norm = c_max(dua_res_norm, t);
  t = vec_norm_inf(w->Px, n); dua_res_norm = c_max(dua_res_norm, t);
  dua_res /= (dua_res_norm + 1e-10);
  double rho_estimate = w->st.rho * sqrt(pri_res / (dua_res + 1e-10));
  rho_estimate = c_min(c_max(rho_estimate, RHO_MIN), RHO_MAX);
  return rho_estimate;
}
static int update_rho(work_t *w, double rho_new) {
  w->st.rho = c_min(c_max(rho_new, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if (w->constr_type[i] == 0) { w->rho_vec[i] = w->st.rho; w->rho_inv_vec[i] = 1. / w->st.rho; }
    else if (w->constr_type[i] == 1) { w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho; w->rho_inv_vec[i] = 1. / w->rho_vec[i]; }
  }
  w->n_factor++;
  return linsys_refactor(w->ls, w->P, w->A, w->st.sigma, w->rho_inv_vec);
}
static int adapt_rho(work_t *w) {
  double rho_new = compute_rho_estimate(w);
  w->rho_estimate = rho_new;
  if ((rho_new > w->st.rho * w->st.adaptive_rho_tolerance) || (rho_new < w->st.rho / w->st.adaptive_rho_tolerance)) {
    int e = update_rho(w, rho_new);
    w->rho_updates += 1;
    return e;
  }
  return 0;
}

static int has_solution(int status) {
  return (status != OSQP_REF_PRIMAL_INFEASIBLE) && (status != OSQP_REF_PRIMAL_INFEASIBLE_INACCURATE) &&
         (status != OSQP_REF_DUAL_INFEASIBLE) && (status != OSQP_REF_DUAL_INFEASIBLE_INACCURATE) &&
         (status != OSQP_REF_NON_CVX);
}

/* ------------------------------------------------------------------ polish */
static int polish(work_t *w, unsigned char *act_lo, unsigned char *act_up) {
  int n = w->n, m = w->m;
  int *A_to_Alow = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  int *A_to_Aupp = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  int *ind_low = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  int *ind_upp = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  int n_low = 0, n_upp = 0;
  for (int j = 0; j < m; j++) {
    if (w->z[j] - w->l[j] < -w->y[j]) { ind_low[n_low] = j; A_to_Alow[j] = n_low++; } else A_to_Alow[j] = -1;
  }
  for (int j = 0; j < m; j++) {
    if (w->u[j] - w->z[j] < w->y[j]) { ind_upp[n_upp] = j; A_to_Aupp[j] = n_upp++; } else A_to_Aupp[j] = -1;
  }
  if (act_lo) for (int j = 0; j < m; j++) act_lo[j] = A_to_Alow[j] != -1;
  if (act_up) for (int j = 0; j < m; j++) act_up[j] = A_to_Aupp[j] != -1;
  int mred = n_low + n_upp;
  int anz = 0;
  for (int p = 0; p < w->A->p[n]; p++)
    if (A_to_Alow[w->A->i[p]] != -1 || A_to_Aupp[w->A->i[p]] != -1) anz++;
  csc *Ared = csc_alloc(mred, n, anz);
  anz = 0;
  for (int j = 0; j < n; j++) {
    Ared->p[j] = anz;
    for (int p = w->A->p[j]; p < w->A->p[j + 1]; p++) {
      int r = w->A->i[p];
      if (A_to_Alow[r] != -1) { Ared->i[anz] = A_to_Alow[r]; Ared->x[anz++] = w->A->x[p]; }
      else if (A_to_Aupp[r] != -1) { Ared->i[anz] = A_to_Aupp[r] + n_low; Ared->x[anz++] = w->A->x[p]; }
    }
  }
  Ared->p[n] = anz;
  double *dvecp = dvec(mred);
  for (int i = 0; i < mred; i++) dvecp[i] = w->st.delta;
  linsys *pl = linsys_init(w->P, Ared, w->st.delta, dvecp, w->st.linsys, 0);
  int ret = 0;
  if (!pl) { w->status_polish = -1; ret = 1; goto done; }
  {
    int dim = n + mred;
    double *rhs_red = dvec(dim), *pol_sol = dvec(dim), *rhs = dvec(dim);
    double *px = dvec(n), *pz = dvec(m), *py = dvec(m);
    for (int j = 0; j < n; j++) rhs_red[j] = -w->q[j];
    for (int j = 0; j < n_low; j++) rhs_red[n + j] = w->l[ind_low[j]];
    for (int j = 0; j < n_upp; j++) rhs_red[n + n_low + j] = w->u[ind_upp[j]];
    memcpy(pol_sol, rhs_red, sizeof(double) * (size_t)dim);
    linsys_solve_raw(pl, pol_sol);
    for (int it = 0; it < w->st.polish_refine_iter; it++) {
      memcpy(rhs, rhs_red, sizeof(double) * (size_t)dim);
      mat_vec(w->P, pol_sol, rhs, -1);
      mat_tpose_vec(w->P, pol_sol, rhs, -1, 1);
      mat_tpose_vec(Ared, pol_sol + n, rhs, -1, 0);
      mat_vec(Ared, pol_sol, rhs + n, -1);
      linsys_solve_raw(pl, rhs);
      for (int j = 0; j < dim; j++) pol_sol[j] += rhs[j];
    }
    memcpy(px, pol_sol, sizeof(double) * (size_t)n);
    mat_vec(w->A, px, pz, 0);
    for (int j = 0; j < m; j++) {
      if (A_to_Alow[j] != -1) py[j] = pol_sol[n + A_to_Alow[j]];
      else if (A_to_Aupp[j] != -1) py[j] = pol_sol[n + A_to_Aupp[j] + n_low];
      else py[j] = 0.0;
    }
    /* project (z,y) onto the normal cone */
    for (int i = 0; i < m; i++) {
      w->z_prev[i] = pz[i] + py[i];
      pz[i] = c_min(c_max(w->z_prev[i], w->l[i]), w->u[i]);
      py[i] = w->z_prev[i] - pz[i];
    }
    double pol_obj = compute_obj_val(w, px);
    double pol_pri = m == 0 ? 0. : compute_pri_res(w, px, pz);
    double pol_dua = compute_dua_res(w, px, py);
    int ok = (pol_pri < w->pri_res && pol_dua < w->dua_res) || (pol_pri < w->pri_res && w->dua_res < 1e-10) ||
             (pol_dua < w->dua_res && w->pri_res < 1e-10);
    if (ok) {
      w->obj_val = pol_obj; w->pri_res = pol_pri; w->dua_res = pol_dua; w->status_polish = 1;
      memcpy(w->x, px, sizeof(double) * (size_t)n);
      memcpy(w->z, pz, sizeof(double) * (size_t)m);
      memcpy(w->y, py, sizeof(double) * (size_t)m);
    } else w->status_polish = -1;
    free(rhs_red); free(pol_sol); free(rhs); free(px); free(pz); free(py);
    linsys_free(pl);
  }
done:
  csc_free(Ared); free(dvecp);
  free(A_to_Alow); free(A_to_Aupp); free(ind_low); free(ind_upp);
  return ret;
}

/* ------------------------------------------------------------------ public */
void osqp_ref_default_settings(osqp_ref_settings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->delta = 1e-6; s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->check_termination = 25; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 0;
  s->polish = 0; s->polish_refine_iter = 3; s->scaled_termination = 0;
  s->linsys = 0; s->cache_ordering = 1;
}

int osqp_ref_solve(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                   const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                   const osqp_ref_settings *settings, osqp_ref_result *res) {
  /* data validation (upstream validate_data): l <= u, P upper triangular */
  for (int i = 0; i < m; i++) if (l[i] > u[i]) return OSQP_REF_DATA_VALIDATION_ERROR;
  for (int j = 0; j < n; j++) for (int p = Pp[j]; p < Pp[j + 1]; p++) if (Pi[p] > j) return OSQP_REF_DATA_VALIDATION_ERROR;

  work_t W; memset(&W, 0, sizeof(W));
  work_t *w = &W;
  w->n = n; w->m = m; w->st = *settings;
  w->P = csc_copy(n, n, Pp, Pi, Px);
  w->A = csc_copy(m, n, Ap, Ai, Ax);
  w->q = dvec(n); memcpy(w->q, q, sizeof(double) * (size_t)n);
  w->l = dvec(m); w->u = dvec(m);
  /* python wrapper: clip infinities to +-OSQP_INFTY */
  for (int i = 0; i < m; i++) { w->l[i] = c_max(l[i], -OSQP_INFTY); w->u[i] = c_min(u[i], OSQP_INFTY); }
  w->D = dvec(n); w->Dinv = dvec(n); w->E = dvec(m); w->Einv = dvec(m);
  w->rho_vec = dvec(m); w->rho_inv_vec = dvec(m); w->constr_type = (int *)calloc((size_t)(m > 0 ? m : 1), sizeof(int));
  w->x = dvec(n); w->y = dvec(m); w->z = dvec(m); w->xz_tilde = dvec(n + m); w->x_prev = dvec(n); w->z_prev = dvec(m);
  w->Ax = dvec(m); w->Px = dvec(n); w->Aty = dvec(n); w->delta_y = dvec(m); w->Atdelta_y = dvec(n);
  w->delta_x = dvec(n); w->Pdelta_x = dvec(n); w->Adelta_x = dvec(m);
  w->D_temp = dvec(n); w->D_temp_A = dvec(n); w->E_temp = dvec(m); w->sol = dvec(n + m);
  w->c = 1.0; w->cinv = 1.0;
  for (int i = 0; i < n; i++) { w->D[i] = 1.; w->Dinv[i] = 1.; }
  for (int i = 0; i < m; i++) { w->E[i] = 1.; w->Einv[i] = 1.; }

  if (w->st.scaling) scale_data(w);
  set_rho_vec(w);
  w->status = OSQP_REF_UNSOLVED;
  w->ls = linsys_init(w->P, w->A, w->st.sigma, w->rho_inv_vec, w->st.linsys, w->st.cache_ordering);
  w->n_factor = 1;
  int ret = 0;
  if (!w->ls) { ret = OSQP_REF_LINSYS_ERROR; goto cleanup; }

  {
    int iter, can_check_termination = 0;
    int max_iter = w->st.max_iter;
    double alpha = w->st.alpha;
    for (iter = 1; iter <= max_iter; iter++) {
      double *t;
      t = w->x; w->x = w->x_prev; w->x_prev = t;
      t = w->z; w->z = w->z_prev; w->z_prev = t;
      /* x_tilde, z_tilde */
      for (int i = 0; i < n; i++) w->xz_tilde[i] = w->st.sigma * w->x_prev[i] - w->q[i];
      for (int i = 0; i < m; i++) w->xz_tilde[i + n] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
      kkt_solve(w, w->xz_tilde);
      /* x */
      for (int i = 0; i < n; i++) w->x[i] = alpha * w->xz_tilde[i] + (1.0 - alpha) * w->x_prev[i];
      for (int i = 0; i < n; i++) w->delta_x[i] = w->x[i] - w->x_prev[i];
      /* z */
      for (int i = 0; i < m; i++)
        w->z[i] = alpha * w->xz_tilde[i + n] + (1.0 - alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
      for (int i = 0; i < m; i++) w->z[i] = c_min(c_max(w->z[i], w->l[i]), w->u[i]);
      /* y */
      for (int i = 0; i < m; i++) {
        w->delta_y[i] = w->rho_vec[i] * (alpha * w->xz_tilde[i + n] + (1.0 - alpha) * w->z_prev[i] - w->z[i]);
        w->y[i] += w->delta_y[i];
      }
      can_check_termination = w->st.check_termination && (iter % w->st.check_termination == 0);
      if (can_check_termination) {
        update_info(w, iter);
        if (check_termination(w, 0)) break;
      }
      if (w->st.adaptive_rho && !w->st.adaptive_rho_interval) {
        if (w->st.check_termination) w->st.adaptive_rho_interval = ADAPTIVE_RHO_MULTIPLE_TERMINATION * w->st.check_termination;
        else w->st.adaptive_rho_interval = ADAPTIVE_RHO_FIXED;
      }
      if (w->st.adaptive_rho && w->st.adaptive_rho_interval && (iter % w->st.adaptive_rho_interval == 0)) {
        if (!can_check_termination) update_info(w, iter);
        if (adapt_rho(w)) { ret = OSQP_REF_LINSYS_ERROR; goto cleanup; }
      }
    }
    if (!can_check_termination) {
      update_info(w, iter - 1);
      check_termination(w, 0);
    }
    if (has_solution(w->status) && w->status != OSQP_REF_NON_CVX) w->obj_val = compute_obj_val(w, w->x);
    if (w->status == OSQP_REF_UNSOLVED) {
      if (!check_termination(w, 1)) w->status = OSQP_REF_MAX_ITER_REACHED;
    }
    w->rho_estimate = compute_rho_estimate(w);
  }

  if (res->xs) memcpy(res->xs, w->x, sizeof(double) * (size_t)n);
  if (res->zs) memcpy(res->zs, w->z, sizeof(double) * (size_t)m);
  if (res->ys) memcpy(res->ys, w->y, sizeof(double) * (size_t)m);
  if (res->active_lo) memset(res->active_lo, 0, (size_t)m);
  if (res->active_up) memset(res->active_up, 0, (size_t)m);
  w->status_polish = 0;
  if (w->st.polish && w->status == OSQP_REF_SOLVED) polish(w, res->active_lo, res->active_up);

  /* store solution (unscaled) */
  if (has_solution(w->status)) {
    for (int i = 0; i < n; i++) res->x[i] = w->st.scaling ? w->D[i] * w->x[i] : w->x[i];
    for (int i = 0; i < m; i++) res->y[i] = w->st.scaling ? w->cinv * (w->E[i] * w->y[i]) : w->y[i];
    if (res->z) for (int i = 0; i < m; i++) res->z[i] = w->st.scaling ? w->Einv[i] * w->z[i] : w->z[i];
  } else {
    for (int i = 0; i < n; i++) res->x[i] = NAN;
    for (int i = 0; i < m; i++) res->y[i] = NAN;
    if (res->z) for (int i = 0; i < m; i++) res->z[i] = NAN;
  }
  res->status = w->status; res->iter = w->iter; res->rho_updates = w->rho_updates;
  res->status_polish = w->status_polish; res->obj_val = w->obj_val; res->pri_res = w->pri_res;
  res->dua_res = w->dua_res; res->rho_estimate = w->rho_estimate; res->n_factor = w->n_factor;
  res->rho_final = w->st.rho; res->c = w->c;
  if (res->D) memcpy(res->D, w->D, sizeof(double) * (size_t)n);
  if (res->E) memcpy(res->E, w->E, sizeof(double) * (size_t)m);
  if (res->Ps) memcpy(res->Ps, w->P->x, sizeof(double) * (size_t)Pp[n]);
  if (res->As) memcpy(res->As, w->A->x, sizeof(double) * (size_t)Ap[n]);
  if (res->qs) memcpy(res->qs, w->q, sizeof(double) * (size_t)n);

cleanup:
  linsys_free(w->ls);
  csc_free(w->P); csc_free(w->A);
  free(w->q); free(w->l); free(w->u); free(w->D); free(w->Dinv); free(w->E); free(w->Einv);
  free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
  free(w->x); free(w->y); free(w->z); free(w->xz_tilde); free(w->x_prev); free(w->z_prev);
  free(w->Ax); free(w->Px); free(w->Aty); free(w->delta_y); free(w->Atdelta_y);
  free(w->delta_x); free(w->Pdelta_x); free(w->Adelta_x);
  free(w->D_temp); free(w->D_temp_A); free(w->E_temp); free(w->sol);
  return ret;
}
