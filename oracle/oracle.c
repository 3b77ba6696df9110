/*
 * oracle.c -- CPU restatement (plain C, fp64) of Neko-TOP's adjoint-RHS hot path.
 * TEST INFRASTRUCTURE ONLY -- see oracle.h for the rules and the "parity unpinned" statement.
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 * Neko's own operators (opgrad, cdtp, conv1, tnsr3d, speclib) are NOT in the reference tree
 * (un-vendored `develop` dependency, scripts/dependencies.sh:181-187); they are restated from
 * their published definitions (Nek5000 speclib / Neko operators module) as used at the
 * reference call sites listed in SURVEY.md section 2.2.
 *
 * Element loops carry `#pragma omp parallel for` so the same code is the timed CPU baseline
 * (OMP_NUM_THREADS=1 == the reference's "1 MPI rank"; all cores == one rank per core).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py sets the thread count explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ============================ speclib ======================================================= */

/* Legendre polynomial P_n(x) and derivative by the three-term recurrence. */
static void legendre(double x, int n, double *p, double *dp) {
  double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
  if (n == 0) { *p = 1.0; *dp = 0.0; return; }
  for (int k = 2; k <= n; k++) {
    double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    double dk = d0 + (2.0 * k - 1.0) * p1;
    p0 = p1; p1 = pk; d0 = d1; d1 = dk;
  }
  *p = p1; *dp = d1;
}

/* zwgll: GLL nodes = {-1, roots of P'_N, +1}, N = np-1; w_i = 2 / (N (N+1) P_N(z_i)^2). */
void orc_zwgll(double *z, double *w, int np) {
  int N = np - 1;
  if (np == 1) { z[0] = 0.0; w[0] = 2.0; return; }
  z[0] = -1.0; z[N] = 1.0;
  for (int i = 1; i < N; i++) {
    double x = -cos(M_PI * i / N);   /* Chebyshev-Gauss-Lobatto start */
    for (int it = 0; it < 100; it++) {
      double p, dp; legendre(x, N, &p, &dp);
      /* q(x) = P'_N ; q'(x) from Legendre ODE: (1-x^2) P'' = 2x P' - N(N+1) P */
      double ddp = (2.0 * x * dp - N * (N + 1.0) * p) / (1.0 - x * x);
      double dx = dp / ddp;
      x -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    z[i] = x;
  }
  /* enforce exact antisymmetry of the node set */
  for (int i = 0; i < np / 2; i++) { double a = 0.5 * (z[N - i] - z[i]); z[i] = -a; z[N - i] = a; }
  if (np % 2) z[N / 2] = 0.0;
  for (int i = 0; i < np; i++) {
    double p, dp; legendre(z[i], N, &p, &dp);
    w[i] = 2.0 / (N * (N + 1.0) * p * p);
  }
}

/* zwgl: Gauss-Legendre nodes = roots of P_np; w_i = 2 / ((1-z^2) P'_np(z)^2). */
void orc_zwgl(double *z, double *w, int np) {
  for (int i = 0; i < np; i++) {
    double x = -cos(M_PI * (i + 0.75) / (np + 0.5));
    for (int it = 0; it < 100; it++) {
      double p, dp; legendre(x, np, &p, &dp);
      double dx = p / dp;
      x -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    z[i] = x;
  }
  for (int i = 0; i < np / 2; i++) { double a = 0.5 * (z[np - 1 - i] - z[i]); z[i] = -a; z[np - 1 - i] = a; }
  if (np % 2) z[np / 2] = 0.0;
  for (int i = 0; i < np; i++) {
    double p, dp; legendre(z[i], np, &p, &dp);
    w[i] = 2.0 / ((1.0 - z[i] * z[i]) * dp * dp);
  }
}

/* dgll: D(i,j) = P_N(z_i) / (P_N(z_j) (z_i - z_j)), D(0,0) = -N(N+1)/4, D(N,N) = +N(N+1)/4. */
void orc_dgll(double *D, const double *z, int np) {
  int N = np - 1;
  for (int j = 0; j < np; j++)
    for (int i = 0; i < np; i++) {
      double v;
      if (i != j) {
        double pi, pj, d; legendre(z[i], N, &pi, &d); legendre(z[j], N, &pj, &d);
        v = pi / (pj * (z[i] - z[j]));
      } else if (i == 0) v = -N * (N + 1.0) / 4.0;
      else if (i == N) v = N * (N + 1.0) / 4.0;
      else v = 0.0;
      D[i + np * j] = v;
    }
}

/* barycentric weights of a node set */
static void bary_weights(double *bw, const double *z, int n) {
  for (int j = 0; j < n; j++) {
    double p = 1.0;
    for (int k = 0; k < n; k++) if (k != j) p *= (z[j] - z[k]);
    bw[j] = 1.0 / p;
  }
}

/* derivative matrix of the Lagrange interpolants on arbitrary nodes (Neko builds the GL-space
 * dx with setup_intp(..., derivative order 1) = fd_weights_full; same matrix). */
void orc_deriv_matrix(double *D, const double *z, int np) {
  double *bw = (double *)malloc(sizeof(double) * np);
  bary_weights(bw, z, np);
  for (int i = 0; i < np; i++) {
    double s = 0.0;
    for (int j = 0; j < np; j++) if (j != i) {
      double v = (bw[j] / bw[i]) / (z[i] - z[j]);
      D[i + np * j] = v; s += v;
    }
    D[i + np * i] = -s;
  }
  free(bw);
}

/* J(a,m) = l_m(zto_a): Lagrange interpolation matrix (Neko interpolator_t / setup_intp order 0) */
void orc_interp_matrix(double *J, const double *zto, int nto, const double *zfrom, int nfrom) {
  double *bw = (double *)malloc(sizeof(double) * nfrom);
  bary_weights(bw, zfrom, nfrom);
  for (int a = 0; a < nto; a++) {
    int hit = -1;
    for (int m = 0; m < nfrom; m++) if (zto[a] == zfrom[m]) hit = m;
    if (hit >= 0) {
      for (int m = 0; m < nfrom; m++) J[a + nto * m] = (m == hit) ? 1.0 : 0.0;
      continue;
    }
    double den = 0.0;
    for (int m = 0; m < nfrom; m++) den += bw[m] / (zto[a] - zfrom[m]);
    for (int m = 0; m < nfrom; m++) J[a + nto * m] = (bw[m] / (zto[a] - zfrom[m])) / den;
  }
  free(bw);
}

/* ============================ local tensor kernels ========================================= */

/* ur,us,ut of one element: ur = D_r u, us = D_s u, ut = D_t u  (D col-major D[i+lx*m]) */
static void local_grad(double *ur, double *us, double *ut, const double *u, const double *D, int lx) {
  for (int k = 0; k < lx; k++)
    for (int j = 0; j < lx; j++)
      for (int i = 0; i < lx; i++) {
        double r = 0.0, s = 0.0, t = 0.0;
        for (int m = 0; m < lx; m++) {
          r += D[i + lx * m] * u[m + lx * (j + lx * k)];
          s += D[j + lx * m] * u[i + lx * (m + lx * k)];
          t += D[k + lx * m] * u[i + lx * (j + lx * m)];
        }
        int p = i + lx * (j + lx * k);
        ur[p] = r; us[p] = s; ut[p] = t;
      }
}

/* out += D_r^T a + D_s^T b + D_t^T c for one element (dxt(i,m) = D(m,i)) */
static void local_gradT(double *out, const double *a, const double *b, const double *c,
                        const double *D, int lx) {
  for (int k = 0; k < lx; k++)
    for (int j = 0; j < lx; j++)
      for (int i = 0; i < lx; i++) {
        double r = 0.0, s = 0.0, t = 0.0;
        for (int m = 0; m < lx; m++) {
          r += D[m + lx * i] * a[m + lx * (j + lx * k)];
          s += D[m + lx * j] * b[i + lx * (m + lx * k)];
          t += D[m + lx * k] * c[i + lx * (j + lx * m)];
        }
        /* Neko's cdtp accumulates direction by direction: dtx = r-part; dtx += s-part; += t-part */
        out[i + lx * (j + lx * k)] = (r + s) + t;
      }
}

static void make_w3(double *w3, const double *w, int lx) {
  for (int k = 0; k < lx; k++)
    for (int j = 0; j < lx; j++)
      for (int i = 0; i < lx; i++) w3[i + lx * (j + lx * k)] = w[i] * w[j] * w[k]; /* space.f90 */
}

/* ============================ Neko operators =============================================== */

/* one element of tnsr3d; t1 (nv*nu*nu) and t2 (nv*nv*nu) are caller scratch */
static void local_tnsr3d(double *ve, int nv, const double *ue, int nu, const double *A,
                         double *t1, double *t2) {
  for (int n = 0; n < nu; n++) for (int m = 0; m < nu; m++) for (int a = 0; a < nv; a++) {
    double s = 0.0; for (int l = 0; l < nu; l++) s += A[a + nv * l] * ue[l + nu * (m + nu * n)];
    t1[a + nv * (m + nu * n)] = s;
  }
  for (int n = 0; n < nu; n++) for (int b = 0; b < nv; b++) for (int a = 0; a < nv; a++) {
    double s = 0.0; for (int m = 0; m < nu; m++) s += t1[a + nv * (m + nu * n)] * A[b + nv * m];
    t2[a + nv * (b + nv * n)] = s;
  }
  for (int c = 0; c < nv; c++) for (int b = 0; b < nv; b++) for (int a = 0; a < nv; a++) {
    double s = 0.0; for (int n = 0; n < nu; n++) s += t2[a + nv * (b + nv * n)] * A[c + nv * n];
    ve[a + nv * (b + nv * c)] = s;
  }
}

void orc_tnsr3d(double *v, int nv, const double *u, int nu, const double *A, int nelv) {
  size_t nu3 = (size_t)nu * nu * nu, nv3 = (size_t)nv * nv * nv;
  int nm = nv > nu ? nv : nu;
#pragma omp parallel
  {
    double *t1 = (double *)malloc(sizeof(double) * 2 * nm * nm * nm), *t2 = t1 + (size_t)nm * nm * nm;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) local_tnsr3d(v + nv3 * e, nv, u + nu3 * e, nu, A, t1, t2);
    free(t1);
  }
}

void orc_tnsr3d_t(double *u, int nu, const double *v, int nv, const double *A, int nelv) {
  /* u = (A^T x A^T x A^T) v : build At (nu x nv) and reuse the forward kernel */
  double *At = (double *)malloc(sizeof(double) * nu * nv);
  for (int a = 0; a < nv; a++) for (int l = 0; l < nu; l++) At[l + nu * a] = A[a + nv * l];
  orc_tnsr3d(u, nu, v, nv, At, nelv);
  free(At);
}

/* opgrad (weak gradient): ux = w3*(drdx*ur + dsdx*us + dtdx*ut), ... */
void orc_opgrad(double *ux, double *uy, double *uz, const double *u, int lx, int nelv,
                const double *D, const double *w3, double *const G[9]) {
  size_t N = (size_t)lx * lx * lx;
#pragma omp parallel
  {
    double *ur = (double *)malloc(sizeof(double) * 3 * N), *us = ur + N, *ut = us + N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      local_grad(ur, us, ut, u + o, D, lx);
      for (size_t i = 0; i < N; i++) {
        ux[o + i] = w3[i] * (G[0][o + i] * ur[i] + G[1][o + i] * us[i] + G[2][o + i] * ut[i]);
        uy[o + i] = w3[i] * (G[3][o + i] * ur[i] + G[4][o + i] * us[i] + G[5][o + i] * ut[i]);
        uz[o + i] = w3[i] * (G[6][o + i] * ur[i] + G[7][o + i] * us[i] + G[8][o + i] * ut[i]);
      }
    }
    free(ur);
  }
}

/* cdtp: dtx = D_r^T(w3*x*dr) + D_s^T(w3*x*ds) + D_t^T(w3*x*dt) */
void orc_cdtp(double *dtx, const double *x, const double *dr, const double *ds, const double *dt,
              int lx, int nelv, const double *D, const double *w3) {
  size_t N = (size_t)lx * lx * lx;
#pragma omp parallel
  {
    double *a = (double *)malloc(sizeof(double) * 3 * N), *b = a + N, *c = b + N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      for (size_t i = 0; i < N; i++) {
        double wx = x[o + i] * w3[i];
        a[i] = wx * dr[o + i]; b[i] = wx * ds[o + i]; c[i] = wx * dt[o + i];
      }
      local_gradT(dtx + o, a, b, c, D, lx);
    }
    free(a);
  }
}

/* conv1: du = jacinv*(vx*(drdx ur+dsdx us+dtdx ut) + vy*(..dy..) + vz*(..dz..)) */
void orc_conv1(double *du, const double *u, const double *vx, const double *vy, const double *vz,
               int lx, int nelv, const double *D, double *const G[9], const double *jacinv) {
  size_t N = (size_t)lx * lx * lx;
#pragma omp parallel
  {
    double *ur = (double *)malloc(sizeof(double) * 3 * N), *us = ur + N, *ut = us + N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      local_grad(ur, us, ut, u + o, D, lx);
      for (size_t i = 0; i < N; i++) {
        du[o + i] = jacinv[o + i] *
            (vx[o + i] * (G[0][o + i] * ur[i] + G[1][o + i] * us[i] + G[2][o + i] * ut[i]) +
             vy[o + i] * (G[3][o + i] * ur[i] + G[4][o + i] * us[i] + G[5][o + i] * ut[i]) +
             vz[o + i] * (G[6][o + i] * ur[i] + G[7][o + i] * us[i] + G[8][o + i] * ut[i]));
      }
    }
    free(ur);
  }
}

/* coef_t: geometric factors (cofactors, J-scaled), jac, B = jac*w3 */
void orc_geom(int lx, int nelv, const double *D, const double *w,
              const double *x, const double *y, const double *z,
              double *G[9], double *jac, double *B) {
  size_t N = (size_t)lx * lx * lx;
  double *w3 = (double *)malloc(sizeof(double) * N);
  make_w3(w3, w, lx);
#pragma omp parallel
  {
    double *t = (double *)malloc(sizeof(double) * 9 * N);
    double *xr = t, *xs = t + N, *xt = t + 2 * N, *yr = t + 3 * N, *ys = t + 4 * N, *yt = t + 5 * N,
           *zr = t + 6 * N, *zs = t + 7 * N, *zt = t + 8 * N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      local_grad(xr, xs, xt, x + o, D, lx);
      local_grad(yr, ys, yt, y + o, D, lx);
      local_grad(zr, zs, zt, z + o, D, lx);
      for (size_t i = 0; i < N; i++) {
        double J = xr[i] * ys[i] * zt[i] + xt[i] * yr[i] * zs[i] + xs[i] * yt[i] * zr[i]
                 - xr[i] * yt[i] * zs[i] - xs[i] * yr[i] * zt[i] - xt[i] * ys[i] * zr[i];
        G[0][o + i] = ys[i] * zt[i] - yt[i] * zs[i];  /* drdx */
        G[1][o + i] = yt[i] * zr[i] - yr[i] * zt[i];  /* dsdx */
        G[2][o + i] = yr[i] * zs[i] - ys[i] * zr[i];  /* dtdx */
        G[3][o + i] = xt[i] * zs[i] - xs[i] * zt[i];  /* drdy */
        G[4][o + i] = xr[i] * zt[i] - xt[i] * zr[i];  /* dsdy */
        G[5][o + i] = xs[i] * zr[i] - xr[i] * zs[i];  /* dtdy */
        G[6][o + i] = xs[i] * yt[i] - xt[i] * ys[i];  /* drdz */
        G[7][o + i] = xt[i] * yr[i] - xr[i] * yt[i];  /* dsdz */
        G[8][o + i] = xr[i] * ys[i] - xs[i] * yr[i];  /* dtdz */
        if (jac) jac[o + i] = J;
        if (B) B[o + i] = J * w3[i];
      }
    }
    free(t);
  }
  free(w3);
}

/* ============================ adjoint advection, no dealias ================================ */

/* One element of the intended operator (device branch adv_adjoint_no_dealias.f90:162-201 and
 * adjoint_weak_no_dealias_device :269-303), arranged element-by-element like the CPU branch
 * (:205-251, :317-346) so it is also the timed CPU baseline. */
static void adj_adv_element(double *fx, double *fy, double *fz,
                            const double *vx, const double *vy, const double *vz,
                            const double *ub, const double *vb, const double *wb,
                            const double *const Ge[9], const double *D, const double *w3, int lx,
                            double *scr, int bug_compat, size_t shift_room) {
  size_t N = (size_t)lx * lx * lx;
  double *ur = scr, *us = scr + N, *ut = scr + 2 * N;
  double *du[9];   /* duxb,duyb,duzb, dvxb,dvyb,dvzb, dwxb,dwyb,dwzb */
  for (int q = 0; q < 9; q++) du[q] = scr + (3 + q) * N;
  double *wk1 = scr + 12 * N, *wk2 = scr + 13 * N, *wk3 = scr + 14 * N;
  double *w1 = scr + 15 * N, *w2 = scr + 16 * N, *w3o = scr + 17 * N;
  double *ca = scr + 18 * N, *cb = scr + 19 * N, *cc = scr + 20 * N;

  /* :165-167 / :208-210  opgrad of each base-flow component */
  const double *U[3] = {ub, vb, wb};
  for (int c = 0; c < 3; c++) {
    local_grad(ur, us, ut, U[c], D, lx);
    for (size_t i = 0; i < N; i++) {
      du[3 * c + 0][i] = w3[i] * (Ge[0][i] * ur[i] + Ge[1][i] * us[i] + Ge[2][i] * ut[i]);
      du[3 * c + 1][i] = w3[i] * (Ge[3][i] * ur[i] + Ge[4][i] * us[i] + Ge[5][i] * ut[i]);
      du[3 * c + 2][i] = w3[i] * (Ge[6][i] * ur[i] + Ge[7][i] * us[i] + Ge[8][i] * ut[i]);
    }
  }
  /* :171-181 vdot3 + sub2 (device) / :214-230 (CPU; D1 shifts the f index by one) */
  size_t sh = (bug_compat && shift_room) ? 1 : 0;
  for (size_t i = 0; i < N; i++) {
    fx[i + sh] -= (vx[i] * du[0][i] + vy[i] * du[3][i] + vz[i] * du[6][i]);
    fy[i + sh] -= (vx[i] * du[1][i] + vy[i] * du[4][i] + vz[i] * du[7][i]);
    fz[i + sh] -= (vx[i] * du[2][i] + vy[i] * du[5][i] + vz[i] * du[8][i]);
  }
  /* :183-201 three calls of adjoint_weak_no_dealias_* */
  double *F[3] = {fx, fy, fz};
  const double *V[3] = {vx, vy, vz};
  for (int c = 0; c < 3; c++) {
    /* :293-295 / :331-335 outer product */
    for (size_t i = 0; i < N; i++) {
      wk1[i] = V[c][i] * ub[i]; wk2[i] = V[c][i] * vb[i]; wk3[i] = V[c][i] * wb[i];
    }
    /* :297-299 / :338-340 three cdtp */
    const double *wk[3] = {wk1, wk2, wk3};
    double *wo[3] = {w1, w2, w3o};
    for (int k = 0; k < 3; k++) {
      for (size_t i = 0; i < N; i++) {
        double wx = wk[k][i] * w3[i];
        ca[i] = wx * Ge[3 * k + 0][i]; cb[i] = wx * Ge[3 * k + 1][i]; cc[i] = wx * Ge[3 * k + 2][i];
      }
      local_gradT(wo[k], ca, cb, cc, D, lx);
    }
    /* :301-302 add4 + sub2 (device, intended) / :344 (CPU; D2 sign defect) */
    if (bug_compat) for (size_t i = 0; i < N; i++) F[c][i] = F[c][i] - w1[i] + w2[i] + w3o[i];
    else for (size_t i = 0; i < N; i++) F[c][i] -= (w1[i] + w2[i] + w3o[i]);
  }
}

void orc_adjoint_advection_no_dealias(double *fx, double *fy, double *fz,
                                      const double *vx, const double *vy, const double *vz,
                                      const double *vxb, const double *vyb, const double *vzb,
                                      int lx, int nelv, const double *D, const double *w,
                                      double *const G[9], int bug_compat) {
  size_t N = (size_t)lx * lx * lx;
  double *w3 = (double *)malloc(sizeof(double) * N);
  make_w3(w3, w, lx);
  /* bug_compat runs serially: D1 makes consecutive elements overlap by one entry */
#pragma omp parallel if (!bug_compat)
  {
    double *scr = (double *)malloc(sizeof(double) * 21 * N);
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      const double *Ge[9];
      for (int q = 0; q < 9; q++) Ge[q] = G[q] + o;
      adj_adv_element(fx + o, fy + o, fz + o, vx + o, vy + o, vz + o, vxb + o, vyb + o, vzb + o,
                      Ge, D, w3, lx, scr, bug_compat, (size_t)(e < nelv - 1));
    }
    free(scr);
  }
  free(w3);
}

/* adv_adjoint_no_dealias.f90:404-424: f_i -= B*conv1(u'_i; U_b) + B*conv1(U_b,i; u') */
void orc_linear_advection_no_dealias(double *fx, double *fy, double *fz,
                                     const double *vx, const double *vy, const double *vz,
                                     const double *vxb, const double *vyb, const double *vzb,
                                     int lx, int nelv, const double *D, const double *w,
                                     double *const G[9], const double *jac) {
  size_t N = (size_t)lx * lx * lx, n = N * nelv;
  double *w3 = (double *)malloc(sizeof(double) * N);
  make_w3(w3, w, lx);
  double *B = (double *)malloc(sizeof(double) * n), *ji = (double *)malloc(sizeof(double) * n);
  double *tmp = (double *)malloc(sizeof(double) * n);
  for (size_t i = 0; i < n; i++) { B[i] = jac[i] * w3[i % N]; ji[i] = 1.0 / jac[i]; }
  double *F[3] = {fx, fy, fz};
  const double *V[3] = {vx, vy, vz}, *Vb[3] = {vxb, vyb, vzb};
  for (int c = 0; c < 3; c++) {
    orc_conv1(tmp, V[c], vxb, vyb, vzb, lx, nelv, D, G, ji);  /* U_b . grad u'_c */
    for (size_t i = 0; i < n; i++) F[c][i] -= B[i] * tmp[i];  /* subcol3 */
    orc_conv1(tmp, Vb[c], vx, vy, vz, lx, nelv, D, G, ji);    /* u' . grad U_b,c */
    for (size_t i = 0; i < n; i++) F[c][i] -= B[i] * tmp[i];
  }
  free(w3); free(B); free(ji); free(tmp);
}

/* ============================ adjoint advection, dealiased ================================= */

typedef struct {
  int lx, lxd;
  double *zg, *wg, *zd, *wd, *J, *Dd, *w3d;
} dealias_space;

static void dealias_space_init(dealias_space *s, int lx, int lxd) {
  s->lx = lx; s->lxd = lxd;
  s->zg = (double *)malloc(sizeof(double) * lx); s->wg = (double *)malloc(sizeof(double) * lx);
  s->zd = (double *)malloc(sizeof(double) * lxd); s->wd = (double *)malloc(sizeof(double) * lxd);
  s->J = (double *)malloc(sizeof(double) * lx * lxd);
  s->Dd = (double *)malloc(sizeof(double) * lxd * lxd);
  s->w3d = (double *)malloc(sizeof(double) * lxd * lxd * lxd);
  orc_zwgll(s->zg, s->wg, lx);                       /* Xh_GLL */
  orc_zwgl(s->zd, s->wd, lxd);                       /* adv_adjoint_dealias.f90:143 Xh_GL%init(GL,..) */
  orc_interp_matrix(s->J, s->zd, lxd, s->zg, lx);    /* :146 GLL_to_GL%init */
  orc_deriv_matrix(s->Dd, s->zd, lxd);               /* GL-space dx */
  make_w3(s->w3d, s->wd, lxd);
}
static void dealias_space_free(dealias_space *s) {
  free(s->zg); free(s->wg); free(s->zd); free(s->wd); free(s->J); free(s->Dd); free(s->w3d);
}

void orc_adjoint_advection_dealias(double *fx, double *fy, double *fz,
                                   const double *vx, const double *vy, const double *vz,
                                   const double *vxb, const double *vyb, const double *vzb,
                                   int lx, int lxd, int nelv, double *const G[9]) {
  dealias_space sp; dealias_space_init(&sp, lx, lxd);
  size_t N = (size_t)lx * lx * lx, Nd = (size_t)lxd * lxd * lxd;
  /* interpolator_t%map(., Xh_GLL) applies the transpose of the GLL->GL matrix */
  double *Jt = (double *)malloc(sizeof(double) * lx * lxd);
  for (int a = 0; a < lxd; a++) for (int l = 0; l < lx; l++) Jt[l + lx * a] = sp.J[a + lxd * l];
#pragma omp parallel
  {
    /* Gd[9] t[6] du[9] tf[3] ur,us,ut ca,cb,cc + tensor scratch */
    double *buf = (double *)malloc(sizeof(double) * (9 + 6 + 9 + 3 + 3 + 3 + 2) * Nd + sizeof(double) * N);
    double *Gd[9]; for (int q = 0; q < 9; q++) Gd[q] = buf + q * Nd;
    double *t[6]; for (int q = 0; q < 6; q++) t[q] = buf + (9 + q) * Nd;   /* tx,ty,tz,txb,tyb,tzb */
    double *du[9]; for (int q = 0; q < 9; q++) du[q] = buf + (15 + q) * Nd; /* duxb,duyb,duzb,dvxb,.. */
    double *tf[3] = {buf + 24 * Nd, buf + 25 * Nd, buf + 26 * Nd};
    double *ur = buf + 27 * Nd, *us = ur + Nd, *ut = us + Nd;
    double *ca = buf + 30 * Nd, *cb = ca + Nd, *cc = cb + Nd;
    double *t1 = buf + 33 * Nd, *t2 = t1 + Nd;
    double *tmp = buf + 35 * Nd;
    double *F[3] = {fx, fy, fz};
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      /* :153-161 geometric factors interpolated to GL (init_dealias; per element here) */
      for (int q = 0; q < 9; q++) local_tnsr3d(Gd[q], lxd, G[q] + o, lx, sp.J, t1, t2);
      /* :360-367 map base flow and adjoint velocity to GL */
      local_tnsr3d(t[3], lxd, vxb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[4], lxd, vyb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[5], lxd, vzb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[0], lxd, vx + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[1], lxd, vy + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[2], lxd, vz + o, lx, sp.J, t1, t2);
      /* :372-374 opgrad of the base flow on c_GL */
      for (int c = 0; c < 3; c++) {
        local_grad(ur, us, ut, t[3 + c], sp.Dd, lxd);
        for (size_t i = 0; i < Nd; i++) {
          du[3 * c + 0][i] = sp.w3d[i] * (Gd[0][i] * ur[i] + Gd[1][i] * us[i] + Gd[2][i] * ut[i]);
          du[3 * c + 1][i] = sp.w3d[i] * (Gd[3][i] * ur[i] + Gd[4][i] * us[i] + Gd[5][i] * ut[i]);
          du[3 * c + 2][i] = sp.w3d[i] * (Gd[6][i] * ur[i] + Gd[7][i] * us[i] + Gd[8][i] * ut[i]);
        }
      }
      /* :377-381 transpose and multiply */
      for (size_t i = 0; i < Nd; i++) {
        tf[0][i] = t[0][i] * du[0][i] + t[1][i] * du[3][i] + t[2][i] * du[6][i];
        tf[1][i] = t[0][i] * du[1][i] + t[1][i] * du[4][i] + t[2][i] * du[7][i];
        tf[2][i] = t[0][i] * du[2][i] + t[1][i] * du[5][i] + t[2][i] * du[8][i];
      }
      /* :384-392 map back (J^T) and sub2 */
      for (int c = 0; c < 3; c++) {
        local_tnsr3d(tmp, lx, tf[c], lxd, Jt, t1, t2);
        for (size_t i = 0; i < N; i++) F[c][o + i] -= tmp[i];
      }
      /* :394-453 weak-form part, component by component */
      for (int c = 0; c < 3; c++) {
        for (int k = 0; k < 3; k++) {
          /* :395-399 outer product u_c * U_k ; :402-404 cdtp with d./dx_k factors */
          for (size_t i = 0; i < Nd; i++) {
            double wx = (t[c][i] * t[3 + k][i]) * sp.w3d[i];
            ca[i] = wx * Gd[3 * k + 0][i]; cb[i] = wx * Gd[3 * k + 1][i]; cc[i] = wx * Gd[3 * k + 2][i];
          }
          local_gradT(tf[k], ca, cb, cc, sp.Dd, lxd);
        }
        /* :407-409 sum; :412-413 map back and sub2 */
        for (size_t i = 0; i < Nd; i++) tf[0][i] = tf[0][i] + tf[1][i] + tf[2][i];
        local_tnsr3d(tmp, lx, tf[0], lxd, Jt, t1, t2);
        for (size_t i = 0; i < N; i++) F[c][o + i] -= tmp[i];
      }
    }
    free(buf);
  }
  free(Jt);
  dealias_space_free(&sp);
}

void orc_linear_advection_dealias(double *fx, double *fy, double *fz,
                                  const double *vx, const double *vy, const double *vz,
                                  const double *vxb, const double *vyb, const double *vzb,
                                  int lx, int lxd, int nelv, double *const G[9]) {
  dealias_space sp; dealias_space_init(&sp, lx, lxd);
  size_t N = (size_t)lx * lx * lx, Nd = (size_t)lxd * lxd * lxd;
  double *Jt = (double *)malloc(sizeof(double) * lx * lxd);
  for (int a = 0; a < lxd; a++) for (int l = 0; l < lx; l++) Jt[l + lx * a] = sp.J[a + lxd * l];
#pragma omp parallel
  {
    double *buf = (double *)malloc(sizeof(double) * (9 + 6 + 3 + 3 + 3 + 2) * Nd + sizeof(double) * N);
    double *Gd[9]; for (int q = 0; q < 9; q++) Gd[q] = buf + q * Nd;
    double *t[6]; for (int q = 0; q < 6; q++) t[q] = buf + (9 + q) * Nd;
    double *ur = buf + 15 * Nd, *us = ur + Nd, *ut = us + Nd;
    double *vr = buf + 18 * Nd, *vs = vr + Nd, *vt = vs + Nd;
    double *tf[3] = {buf + 21 * Nd, buf + 22 * Nd, buf + 23 * Nd};
    double *t1 = buf + 24 * Nd, *t2 = t1 + Nd;
    double *tmp = buf + 26 * Nd;
    double *F[3] = {fx, fy, fz};
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      for (int q = 0; q < 9; q++) local_tnsr3d(Gd[q], lxd, G[q] + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[3], lxd, vxb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[4], lxd, vyb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[5], lxd, vzb + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[0], lxd, vx + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[1], lxd, vy + o, lx, sp.J, t1, t2);
      local_tnsr3d(t[2], lxd, vz + o, lx, sp.J, t1, t2);
      /* adv_adjoint_dealias.f90:605-664: pass 0 = u'.grad U_b, pass 1 = U_b.grad u' */
      for (int pass = 0; pass < 2; pass++) {
        double **diff = pass == 0 ? &t[3] : &t[0];   /* field being differentiated */
        double **adv = pass == 0 ? &t[0] : &t[3];    /* advecting velocity */
        for (int c = 0; c < 3; c++) {
          local_grad(ur, us, ut, diff[c], sp.Dd, lxd);
          for (size_t i = 0; i < Nd; i++) {
            vr[i] = sp.w3d[i] * (Gd[0][i] * ur[i] + Gd[1][i] * us[i] + Gd[2][i] * ut[i]);
            vs[i] = sp.w3d[i] * (Gd[3][i] * ur[i] + Gd[4][i] * us[i] + Gd[5][i] * ut[i]);
            vt[i] = sp.w3d[i] * (Gd[6][i] * ur[i] + Gd[7][i] * us[i] + Gd[8][i] * ut[i]);
            tf[c][i] = adv[0][i] * vr[i] + adv[1][i] * vs[i] + adv[2][i] * vt[i];
          }
        }
        for (int c = 0; c < 3; c++) {
          local_tnsr3d(tmp, lx, tf[c], lxd, Jt, t1, t2);
          for (size_t i = 0; i < N; i++) F[c][o + i] -= tmp[i];
        }
      }
    }
    free(buf);
  }
  free(Jt);
  dealias_space_free(&sp);
}

/* ============================ minimum-dissipation objective chain ========================== */

void orc_dudxyz(double *du, const double *u, const double *dr, const double *ds, const double *dt,
                const double *jacinv, int lx, int nelv, const double *D) {
  size_t N = (size_t)lx * lx * lx;
#pragma omp parallel
  {
    double *ur = (double *)malloc(sizeof(double) * 3 * N), *us = ur + N, *ut = us + N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      local_grad(ur, us, ut, u + o, D, lx);
      for (size_t i = 0; i < N; i++)
        du[o + i] = jacinv[o + i] * (dr[o + i] * ur[i] + ds[o + i] * us[i] + dt[o + i] * ut[i]);
    }
    free(ur);
  }
}

void orc_curl(double *w1, double *w2, double *w3, const double *u1, const double *u2, const double *u3,
              int lx, int nelv, const double *D, double *const G[9], const double *jacinv,
              const double *B, const double *Binv, const int64_t *class_id, int64_t nclass) {
  size_t n = (size_t)lx * lx * lx * nelv;
  double *a = (double *)malloc(sizeof(double) * n), *b = (double *)malloc(sizeof(double) * n);
  /* G[0..8] = drdx,dsdx,dtdx, drdy,dsdy,dtdy, drdz,dsdz,dtdz */
  orc_dudxyz(a, u3, G[3], G[4], G[5], jacinv, lx, nelv, D);   /* dw/dy */
  orc_dudxyz(b, u2, G[6], G[7], G[8], jacinv, lx, nelv, D);   /* dv/dz */
  for (size_t i = 0; i < n; i++) w1[i] = a[i] - b[i];
  orc_dudxyz(a, u1, G[6], G[7], G[8], jacinv, lx, nelv, D);   /* du/dz */
  orc_dudxyz(b, u3, G[0], G[1], G[2], jacinv, lx, nelv, D);   /* dw/dx */
  for (size_t i = 0; i < n; i++) w2[i] = a[i] - b[i];
  orc_dudxyz(a, u2, G[0], G[1], G[2], jacinv, lx, nelv, D);   /* dv/dx */
  orc_dudxyz(b, u1, G[3], G[4], G[5], jacinv, lx, nelv, D);   /* du/dy */
  for (size_t i = 0; i < n; i++) w3[i] = a[i] - b[i];
  free(a); free(b);
  double *W[3] = {w1, w2, w3};
  for (int c = 0; c < 3; c++) {
    for (size_t i = 0; i < n; i++) W[c][i] *= B[i];
    orc_gs_add(W[c], class_id, nclass, (int64_t)n);
    for (size_t i = 0; i < n; i++) W[c][i] *= Binv[i];
  }
}

void orc_mask_exterior_const(double *fld, const int *mask, int mask_size, double c, int64_t n) {
  double *work = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; i++) work[i] = c;
  for (int m = 0; m < mask_size; m++) work[mask[m] - 1] = fld[mask[m] - 1];
  memcpy(fld, work, sizeof(double) * (size_t)n);
  free(work);
}

double orc_glsc2_mask(const double *a, const double *b, const int *mask, int mask_size, int64_t n) {
  double s = 0.0;
  if (mask) { for (int m = 0; m < mask_size; m++) s += a[mask[m] - 1] * b[mask[m] - 1]; }
  else { for (int64_t i = 0; i < n; i++) s += a[i] * b[i]; }
  return s;
}

void orc_curlcurl_forcing(double *fu, double *fv, double *fw, const double *u, const double *v,
                          const double *w, int lx, int nelv, const double *D, double *const G[9],
                          const double *jacinv, const double *B, const double *Binv,
                          const int64_t *class_id, int64_t nclass, const int *mask, int mask_size,
                          double obj_scale) {
  size_t n = (size_t)lx * lx * lx * nelv;
  double *wo = (double *)malloc(sizeof(double) * 6 * n);
  double *o1 = wo, *o2 = wo + n, *o3 = wo + 2 * n, *o4 = wo + 3 * n, *o5 = wo + 4 * n, *o6 = wo + 5 * n;
  orc_curl(o1, o2, o3, u, v, w, lx, nelv, D, G, jacinv, B, Binv, class_id, nclass);       /* :231 */
  orc_curl(o4, o5, o6, o1, o2, o3, lx, nelv, D, G, jacinv, B, Binv, class_id, nclass);    /* :232 */
  if (mask) {                                                                             /* :235-239 */
    orc_mask_exterior_const(o4, mask, mask_size, 0.0, (int64_t)n);
    orc_mask_exterior_const(o5, mask, mask_size, 0.0, (int64_t)n);
    orc_mask_exterior_const(o6, mask, mask_size, 0.0, (int64_t)n);
  }
  for (size_t i = 0; i < n; i++) {                                                        /* :241-243 */
    fu[i] = fu[i] + obj_scale * o4[i];
    fv[i] = fv[i] + obj_scale * o5[i];
    fw[i] = fw[i] + obj_scale * o6[i];
  }
  free(wo);
}

double orc_min_dissipation_objective(double out[2], const double *u, const double *v, const double *w,
                                     const double *chi, int lx, int nelv, const double *D,
                                     double *const G[9], const double *jacinv, const double *B,
                                     const int *mask, int mask_size, double K, double obj_scale) {
  size_t n = (size_t)lx * lx * lx * nelv;
  double *obj = (double *)calloc(n, sizeof(double)), *g = (double *)malloc(sizeof(double) * n);
  const double *U[3] = {u, v, w};
  for (int c = 0; c < 3; c++)            /* :200-213 grad = dudxyz in x, y, z */
    for (int d = 0; d < 3; d++) {
      orc_dudxyz(g, U[c], G[3 * d], G[3 * d + 1], G[3 * d + 2], jacinv, lx, nelv, D);
      for (size_t i = 0; i < n; i++) obj[i] += g[i] * g[i];
    }
  out[0] = orc_glsc2_mask(obj, B, mask, mask_size, (int64_t)n);                           /* :217-222 */
  out[1] = 0.0;
  if (chi) {   /* :230-239 -- as written: col3(obj,u,chi); addcol3(obj,v,chi); addcol3(obj,w,chi) = (u+v+w)*chi */
    for (size_t i = 0; i < n; i++) obj[i] = u[i] * chi[i];
    for (size_t i = 0; i < n; i++) obj[i] += v[i] * chi[i];
    for (size_t i = 0; i < n; i++) obj[i] += w[i] * chi[i];
    out[1] = orc_glsc2_mask(obj, B, mask, mask_size, (int64_t)n);
  }
  free(obj); free(g);
  return (out[0] + (chi ? 0.5 * K * out[1] : 0.0)) * obj_scale;
}

/* ============================ Helmholtz operator (Neko ax_helm, restated) ================== */

void orc_ax_helm(double *w, const double *u, int lx, int nelv, const double *D, const double *wq,
                 double *const G[9], const double *jacinv, const double *B, double h1, double h2) {
  size_t N = (size_t)lx * lx * lx;
  double *w3 = (double *)malloc(sizeof(double) * N);
  make_w3(w3, wq, lx);
#pragma omp parallel
  {
    double *ur = (double *)malloc(sizeof(double) * 6 * N), *us = ur + N, *ut = us + N;
    double *a = ut + N, *b = a + N, *c = b + N;
#pragma omp for schedule(static)
    for (int e = 0; e < nelv; e++) {
      size_t o = N * e;
      local_grad(ur, us, ut, u + o, D, lx);
      for (size_t i = 0; i < N; i++) {
        const double sc = jacinv[o + i] * w3[i];
        const double rx = G[0][o + i], sx = G[1][o + i], tx = G[2][o + i];
        const double ry = G[3][o + i], sy = G[4][o + i], ty = G[5][o + i];
        const double rz = G[6][o + i], sz = G[7][o + i], tz = G[8][o + i];
        const double G11 = (rx * rx + ry * ry + rz * rz) * sc, G22 = (sx * sx + sy * sy + sz * sz) * sc;
        const double G33 = (tx * tx + ty * ty + tz * tz) * sc, G12 = (rx * sx + ry * sy + rz * sz) * sc;
        const double G13 = (rx * tx + ry * ty + rz * tz) * sc, G23 = (sx * tx + sy * ty + sz * tz) * sc;
        a[i] = h1 * (G11 * ur[i] + G12 * us[i] + G13 * ut[i]);
        b[i] = h1 * (G12 * ur[i] + G22 * us[i] + G23 * ut[i]);
        c[i] = h1 * (G13 * ur[i] + G23 * us[i] + G33 * ut[i]);
      }
      local_gradT(w + o, a, b, c, D, lx);
      for (size_t i = 0; i < N; i++) w[o + i] += h2 * B[o + i] * u[o + i];
    }
    free(ur);
  }
  free(w3);
}

/* ============================ explicit time scheme (Neko rhs_maker, restated) ============== */

void orc_sumab(double *ue, double *ve, double *we, const double *u, const double *v, const double *w,
               const double *ulag1, const double *vlag1, const double *wlag1,
               const double *ulag2, const double *vlag2, const double *wlag2,
               const double ab[3], int nab, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    ue[i] = ab[0] * u[i] + ab[1] * ulag1[i];
    ve[i] = ab[0] * v[i] + ab[1] * vlag1[i];
    we[i] = ab[0] * w[i] + ab[1] * wlag1[i];
  }
  if (nab == 3) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      ue[i] = ue[i] + ab[2] * ulag2[i];
      ve[i] = ve[i] + ab[2] * vlag2[i];
      we[i] = we[i] + ab[2] * wlag2[i];
    }
  }
}

void orc_makeabf(double *abx1, double *aby1, double *abz1, double *abx2, double *aby2, double *abz2,
                 double *fx, double *fy, double *fz, double rho, const double ext[3], int64_t n) {
  double *l1[3] = {abx1, aby1, abz1}, *l2[3] = {abx2, aby2, abz2}, *f[3] = {fx, fy, fz};
  for (int c = 0; c < 3; c++) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      const double ta = ext[1] * l1[c][i] + ext[2] * l2[c][i];
      l2[c][i] = l1[c][i];
      l1[c][i] = f[c][i];
      f[c][i] = (ext[0] * f[c][i] + ta) * rho;
    }
  }
}

void orc_makebdf(const double *ulag1, const double *vlag1, const double *wlag1,
                 const double *ulag2, const double *vlag2, const double *wlag2,
                 double *fx, double *fy, double *fz, const double *u, const double *v, const double *w,
                 const double *B, double rho, double dt, const double bd[4], int nbd, int64_t n) {
  const double *q[3] = {u, v, w}, *l1[3] = {ulag1, vlag1, wlag1}, *l2[3] = {ulag2, vlag2, wlag2};
  double *f[3] = {fx, fy, fz};
  for (int c = 0; c < 3; c++) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      double tb = q[c][i] * B[i] * bd[1];
      if (nbd >= 2) tb = tb + l1[c][i] * B[i] * bd[2];
      if (nbd >= 3) tb = tb + l2[c][i] * B[i] * bd[3];
      f[c][i] = f[c][i] + tb * (rho / dt);
    }
  }
}

/* ============================ pointwise terms ============================================== */

void orc_ramp(double *chi, const double *rho, int64_t n, double f_min, double f_max, double q,
              int convex_up) {
  if (convex_up) {  /* RAMP_mapping.f90:235-239 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) chi[i] = f_min + (f_max - f_min) * rho[i] * (1.0 + q) / (rho[i] + q);
  } else {          /* :190-194 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) chi[i] = f_min + (f_max - f_min) * rho[i] / (1.0 + q * (1.0 - rho[i]));
  }
}

void orc_ramp_backward(double *dF_drho, const double *dF_dchi, const double *rho, int64_t n,
                       double f_min, double f_max, double q, int convex_up) {
  if (convex_up) {  /* :261-265 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
      dF_drho[i] = (f_max - f_min) * (q + 1.0) / ((rho[i] + q) * (rho[i] + q)) * dF_dchi[i];
  } else {          /* :216-220 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      double d = 1.0 - q * (rho[i] - 1.0);
      dF_drho[i] = (f_max - f_min) * (q + 1.0) / (d * d) * dF_dchi[i];
    }
  }
}

void orc_brinkman(double *fx, double *fy, double *fz, const double *u, const double *v,
                  const double *w, const double *chi, int64_t n) {
  /* field_subcol3(f, u, chi): f = f - u*chi   (simple_brinkman_source_term.f90:149-151) */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    fx[i] -= u[i] * chi[i]; fy[i] -= v[i] * chi[i]; fz[i] -= w[i] * chi[i];
  }
}

void orc_lube(double *fx, double *fy, double *fz, const double *u, const double *v,
              const double *w, const double *chi, double K, const int *mask, int mask_size,
              int64_t n) {
  double *work = (double *)malloc(sizeof(double) * n);
  for (int64_t i = 0; i < n; i++) work[i] = chi[i] * K;   /* :190-193 copy + cmult */
  if (mask) {                                               /* :196-198, mask_ops.f90:55-82 */
    double *keep = (double *)calloc(n, sizeof(double));
    for (int m = 0; m < mask_size; m++) keep[mask[m] - 1] = work[mask[m] - 1];
    memcpy(work, keep, sizeof(double) * n); free(keep);
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {                         /* :201-203 addcol3 */
    fx[i] += u[i] * work[i]; fy[i] += v[i] * work[i]; fz[i] += w[i] * work[i];
  }
  free(work);
}

void orc_opcolv(double *fx, double *fy, double *fz, const double *B, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) { fx[i] *= B[i]; fy[i] *= B[i]; fz[i] *= B[i]; }
}

void orc_sensitivity(double *S, const double *u, const double *v, const double *w,
                     const double *ua, const double *va, const double *wa, double K_obj,
                     int if_lube, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double s = u[i] * ua[i];       /* :273 col3    */
    s += v[i] * va[i];             /* :274 addcol3 */
    s += w[i] * wa[i];             /* :275         */
    s *= -1.0;                     /* :277 cmult   */
    if (if_lube) {
      double l = u[i] * u[i];      /* :289-291 */
      l += v[i] * v[i];
      l += w[i] * w[i];
      s += K_obj * l;              /* :293-294 add2s2 */
    }
    S[i] = s;
  }
}

/* ============================ the RHS slice ================================================ */

void orc_adjoint_rhs(double *fx, double *fy, double *fz, double *sens, double *chi_out,
                     const double *vx, const double *vy, const double *vz,
                     const double *vxb, const double *vyb, const double *vzb,
                     const double *rho, const double *chi_in,
                     const double *fsx, const double *fsy, const double *fsz,
                     const int *mask, int mask_size,
                     int lx, int nelv, const double *D, const double *w,
                     double *const G[9], const double *B, const orc_params *p) {
  int64_t n = (int64_t)lx * lx * lx * nelv;
  double *chi = chi_out ? chi_out : (double *)malloc(sizeof(double) * n);
  /* designs/topopt_design.f90:310-311 -> RAMP_mapping.f90:137-153 */
  if (chi_in) memcpy(chi, chi_in, sizeof(double) * n);
  else orc_ramp(chi, rho, n, p->f_min, p->f_max, p->q, p->convex_up);
  /* adjoint_pnpn.f90:669 source_term%compute: handler zeroes f, then sums the terms */
  memset(fx, 0, sizeof(double) * n); memset(fy, 0, sizeof(double) * n); memset(fz, 0, sizeof(double) * n);
  orc_brinkman(fx, fy, fz, vx, vy, vz, chi, n);                 /* steady_state_problem.f90:156-163 */
  if (fsx) for (int64_t i = 0; i < n; i++) { fx[i] += fsx[i]; fy[i] += fsy[i]; fz[i] += fsz[i]; }
  if (p->if_lube) orc_lube(fx, fy, fz, vxb, vyb, vzb, chi, p->K_lube, mask, mask_size, n);
  orc_opcolv(fx, fy, fz, B, n);                                  /* adjoint_pnpn.f90:672-676 */
  if (p->lxd > 0)                                                /* :680-682 adv%compute_adjoint */
    orc_adjoint_advection_dealias(fx, fy, fz, vx, vy, vz, vxb, vyb, vzb, lx, p->lxd, nelv, G);
  else
    orc_adjoint_advection_no_dealias(fx, fy, fz, vx, vy, vz, vxb, vyb, vzb, lx, nelv, D, w, G, 0);
  if (sens) orc_sensitivity(sens, vxb, vyb, vzb, vx, vy, vz, p->K_sens, p->if_lube, n);
  if (!chi_out) free(chi);
}

/* ============================ gather-scatter =============================================== */

typedef struct { int64_t key, idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int64_t orc_gs_classes(int64_t *class_id, const int64_t *key, int64_t n) {
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * n);
  for (int64_t i = 0; i < n; i++) { kv[i].key = key[i]; kv[i].idx = i; }
  qsort(kv, n, sizeof(kv_t), kv_cmp);
  /* representative of each class = its smallest dof index; then relabel by first appearance */
  int64_t *rep = (int64_t *)malloc(sizeof(int64_t) * n);
  for (int64_t i = 0; i < n;) {
    int64_t j = i;
    while (j < n && kv[j].key == kv[i].key) { rep[kv[j].idx] = kv[i].idx; j++; }
    i = j;
  }
  int64_t nclass = 0;
  for (int64_t i = 0; i < n; i++) {
    if (rep[i] == i) class_id[i] = nclass++;
    else class_id[i] = class_id[rep[i]];
  }
  free(kv); free(rep);
  return nclass;
}

void orc_gs_add(double *f, const int64_t *class_id, int64_t nclass, int64_t n) {
  double *acc = (double *)calloc(nclass, sizeof(double));
  for (int64_t i = 0; i < n; i++) acc[class_id[i]] += f[i];   /* ascending dof order */
  for (int64_t i = 0; i < n; i++) f[i] = acc[class_id[i]];
  free(acc);
}
