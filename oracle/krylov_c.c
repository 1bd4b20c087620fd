/* TEST INFRASTRUCTURE (oracle): CPU restatement of what dolfin's solve(A, x, b, "bicgstab", "default") runs in a
 * serial process -- PETSc KSPBCGS with PCILU, i.e. ILU(0) in the natural ordering, left preconditioning, convergence on
 * ||M^-1 r||_2 <= max(rtol ||M^-1 b||_2, atol), nonzero initial guess (reference call sites:
 * lid-driven-cavity/comp_LDC_mesh_convergence.py:425,446,475,495,518,543; PETSc 3.7.5 pinned by fenics-install.sh:28-40;
 * PETSc itself is not vendored in /root/reference, so this follows its published algorithm: KSPBCGS = van der Vorst's
 * BiCGStab with the half-step exit, PCILU levels 0 = IKJ Gaussian elimination restricted to the pattern).
 * Used only by tests/ and bench.py's CPU legs (cpu_baseline / --impl reference); never by the product. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ILU(0) of a CSR matrix with ascending columns; lu gets L (unit diagonal, strictly lower part) and U in A's pattern,
 * dptr[i] = position of the diagonal of row i.  Returns 0, or i+1 when row i has no / a zero pivot. */
int fxo_ilu0(int n, const int* rowptr, const int* col, const double* val, double* lu, int* dptr) {
  int* pos = (int*)malloc(sizeof(int) * (size_t)n);
  if (!pos) return -1;
  for (int i = 0; i < n; ++i) pos[i] = -1;
  memcpy(lu, val, sizeof(double) * (size_t)rowptr[n]);
  for (int i = 0; i < n; ++i) {
    dptr[i] = -1;
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
      pos[col[j]] = j;
      if (col[j] == i) dptr[i] = j;
    }
    if (dptr[i] < 0) { free(pos); return i + 1; }
    for (int j = rowptr[i]; j < dptr[i]; ++j) {
      const int k = col[j];
      const double f = lu[j] / lu[dptr[k]];
      lu[j] = f;
      if (f != 0.0)
        for (int q = dptr[k] + 1; q < rowptr[k + 1]; ++q) {
          const int p = pos[col[q]];
          if (p >= 0) lu[p] -= f * lu[q];
        }
    }
    if (lu[dptr[i]] == 0.0) { free(pos); return i + 1; }
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) pos[col[j]] = -1;
  }
  free(pos);
  return 0;
}

static void ilu_apply(int n, const int* rowptr, const int* col, const double* lu, const int* dptr, const double* r,
                      double* z) {
  for (int i = 0; i < n; ++i) {
    double s = r[i];
    for (int j = rowptr[i]; j < dptr[i]; ++j) s -= lu[j] * z[col[j]];
    z[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int j = dptr[i] + 1; j < rowptr[i + 1]; ++j) s -= lu[j] * z[col[j]];
    z[i] = s / lu[dptr[i]];
  }
}

static void spmv(int n, const int* rowptr, const int* col, const double* val, const double* x, double* y) {
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) s += val[j] * x[col[j]];
    y[i] = s;
  }
}

static double dot(int n, const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* pc: 0 none, 1 Jacobi, 2 ILU(0).  method: 0 BiCGStab, 1 CG.  x: initial guess in, solution out.
 * out[0] = iterations, out[1] = final ||M^-1 r||, out[2] = ||M^-1 b||.  Returns 1 converged, 0 maxit, -1 breakdown,
 * -2 factorisation failed. */
int fxo_krylov(int n, const int* rowptr, const int* col, const double* val, double* x, const double* b, int method,
               int pc, double rtol, double atol, int maxit, double* out) {
  double* w = (double*)calloc((size_t)n * 9, sizeof(double));
  double* lu = NULL;
  int* dptr = NULL;
  int rc = 0, it = 0;
  double rn = 0.0, bn = 0.0;
  if (pc == 2) {
    lu = (double*)malloc(sizeof(double) * (size_t)rowptr[n]);
    dptr = (int*)malloc(sizeof(int) * (size_t)n);
    if (fxo_ilu0(n, rowptr, col, val, lu, dptr) != 0) { rc = -2; goto done; }
  } else if (pc == 1) {
    lu = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {
      lu[i] = 1.0;
      for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
        if (col[j] == i && val[j] != 0.0) lu[i] = 1.0 / val[j];
    }
  }
#define PREC(src, dst)                                                     \
  do {                                                                     \
    if (pc == 2) ilu_apply(n, rowptr, col, lu, dptr, (src), (dst));        \
    else if (pc == 1) for (int i_ = 0; i_ < n; ++i_) (dst)[i_] = lu[i_] * (src)[i_]; \
    else memcpy((dst), (src), sizeof(double) * (size_t)n);                 \
  } while (0)
  {
    double *r = w, *rh = w + n, *p = w + 2 * (size_t)n, *v = w + 3 * (size_t)n, *s = w + 4 * (size_t)n,
           *t = w + 5 * (size_t)n, *tmp = w + 6 * (size_t)n, *z = w + 7 * (size_t)n, *q = w + 8 * (size_t)n;
    PREC(b, tmp);
    bn = sqrt(dot(n, tmp, tmp));
    const double tol = fmax(rtol * bn, atol);
    spmv(n, rowptr, col, val, x, tmp);
    for (int i = 0; i < n; ++i) tmp[i] = b[i] - tmp[i];
    if (method == 0) {
      PREC(tmp, r);
      rn = sqrt(dot(n, r, r));
      if (rn <= tol) { rc = 1; goto done; }
      memcpy(rh, r, sizeof(double) * (size_t)n);
      memcpy(p, r, sizeof(double) * (size_t)n);
      double rho = dot(n, rh, r);
      for (it = 1; it <= maxit; ++it) {
        spmv(n, rowptr, col, val, p, tmp);
        PREC(tmp, v);
        const double d1 = dot(n, rh, v);
        if (d1 == 0.0) { rc = -1; goto done; }
        const double alpha = rho / d1;
        for (int i = 0; i < n; ++i) s[i] = r[i] - alpha * v[i];
        const double sn = sqrt(dot(n, s, s));
        if (sn <= tol) {
          for (int i = 0; i < n; ++i) x[i] += alpha * p[i];
          rn = sn; rc = 1; goto done;
        }
        spmv(n, rowptr, col, val, s, tmp);
        PREC(tmp, t);
        const double tt = dot(n, t, t);
        if (tt == 0.0) { rc = -1; goto done; }
        const double omega = dot(n, t, s) / tt;
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i] + omega * s[i]; r[i] = s[i] - omega * t[i]; }
        rn = sqrt(dot(n, r, r));
        if (rn <= tol) { rc = 1; goto done; }
        if (!(rn == rn)) { rc = -1; goto done; }
        const double rho_new = dot(n, rh, r);
        const double beta = (rho_new / rho) * (alpha / omega);
        for (int i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
        rho = rho_new;
      }
      it = maxit;
    } else { /* preconditioned CG, monitored norm ||M^-1 r|| (KSPCG default) */
      memcpy(r, tmp, sizeof(double) * (size_t)n);
      PREC(r, z);
      rn = sqrt(dot(n, z, z));
      if (rn <= tol) { rc = 1; goto done; }
      memcpy(p, z, sizeof(double) * (size_t)n);
      double gamma = dot(n, r, z);
      for (it = 1; it <= maxit; ++it) {
        spmv(n, rowptr, col, val, p, q);
        const double den = dot(n, p, q);
        if (!(den > 0.0)) { rc = -1; goto done; }
        const double alpha = gamma / den;
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
        PREC(r, z);
        rn = sqrt(dot(n, z, z));
        if (rn <= tol) { rc = 1; goto done; }
        const double gnew = dot(n, r, z);
        const double beta = gnew / gamma;
        for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        gamma = gnew;
      }
      it = maxit;
    }
  }
done:
  out[0] = (double)it; out[1] = rn; out[2] = bn;
  free(w); free(lu); free(dptr);
  return rc;
}
