// Fine-grid advection operators for sm_100a (fp64): the dealiased adjoint operator
// and the linearised operator with and without dealiasing (row f2), one element per CTA pass.
//
// Reference (relative to /root/reference/sources):
//   adjoint/adv_adjoint_dealias.f90:235-462  compute_adjoint_advection_dealias  (MODE = ADV_ADJOINT)
//   adjoint/adv_adjoint_dealias.f90:479-668  compute_linear, dealiased          (MODE = ADV_LINEAR, LXD > LX)
//   adjoint/adv_adjoint_no_dealias.f90:365-427 compute_linear, GLL grid         (MODE = ADV_LINEAR, LXD == LX)
// The reference runs the dealiased operator as ~60 whole-field sweeps over 20 work arrays of nelv*lxd^3
// doubles (:166-208); here the six fine-grid fields of ONE element live in shared memory and nothing but
// the inputs (6 GLL fields + 9 fine-grid geometric factors) and the result ever touches HBM.
//
// Per element (fine grid = Gauss-Legendre lxd^3 when dealiasing, else the GLL grid itself):
//   phase 0  T_q = (J x J x J) field_q, q = v(3), U_b(3)             GLL_to_GL%map         (:360-367)
//   phase 1  thread (i,j) walks its k-column:
//     ADJOINT  dU_c/dx_d = w3 (G_rd d_r + G_sd d_s + G_td d_t) U_c   opgrad on coef_GL     (:372-374)
//              R_d  = sum_c v_c dU_c/dx_d                             vdot3                 (:377-381)
//              R_c += D_r^T(v_c c_r) + D_s^T(v_c c_s) + D_t^T(v_c c_t),
//                     c_r = w3 sum_k U_k G_rk  (contravariant base flow: the 27 transposed contractions of
//                     the 9 cdtp calls, :394-453, collapse to 9 because cdtp is linear in its argument)
//     LINEAR   R_c  = sum_d v_d dU_c/dx_d + sum_d U_d dv_c/dx_d       (:605-664 / no_dealias :404-424,
//                     where B*jacinv == w3)
//   phase 2  f_c -= (J^T x J^T x J^T) R_c                             map(., Xh_GLL) + sub2 (:384-392)
//            (the reference projects six times; projection is linear, so once per component here)
// With FLAG_ACCUM f is in/out (the advection_adjoint_t contract); otherwise the epilogue also evaluates the
// point-wise part of the fused right-hand side (adjrhs_common.cuh header) at the GLL point and f is
// write-only.
//
// Data flow: r/s contractions read shared memory (D rows in registers would not fit beside the 3*LXD
// register accumulators of the column); t contractions read the thread's own column (conflict-free) with
// D(k,m) folded into the instruction as a constant-bank operand (k loop fully unrolled); the transposed
// r/s contractions go through two double-buffered plane buffers (one barrier per plane); the transposed t
// contraction and phase-2's first (k) contraction stay in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "adjrhs_common.cuh"   // FLAG_*
#include "advop.h"           // ADV_ADJOINT / ADV_LINEAR

namespace b200 {

template <int LX, int LXD>
struct AdvParams {
  double D[LXD * LXD];     // fine-grid derivative matrix D(i,m) at D[i + LXD*m]
  double J[LXD * LX];      // GLL -> fine interpolation J(a,l) at J[a + LXD*l] (unused if LXD == LX)
  double wd[LXD];          // fine-grid quadrature weights
  const double* v[3];      // adjoint / perturbation velocity (GLL)
  const double* vb[3];     // base flow (GLL)
  const double* G[9];      // geometric factors on the fine grid, nelv*LXD^3 each
  double* f[3];
  const double* rho;       // rho (FLAG_RAMP) or chi; GLL
  const double* B;         // GLL mass matrix
  const double* fs[3];     // static forcing (FLAG_FSTATIC)
  double* sens;
  double* chi_out;
  const int* elem_list;
  int nelem;
  int elem_base;                // first element of a contiguous range (elem_list == NULL)
  unsigned flags;
  double f_min, f_max, q, K_lube, K_sens;
};

template <int LX, int LXD, int MODE>
struct AdvCfg {
  static constexpr int N = LX * LX * LX, ND = LXD * LXD * LXD, PL = LXD * LXD;
  static constexpr bool INTERP = (LXD != LX);
  static constexpr int NTHR = ((PL + 31) / 32) * 32;
  static constexpr int S1 = LXD * LX * LX, S2 = LXD * LXD * LX;
  static constexpr int NB = 3;                                   // fields interpolated per batch
  static constexpr int X_INTERP = INTERP ? NB * S1 : 0;
  static constexpr int X_PLANES = (MODE == ADV_ADJOINT) ? 2 * 6 * PL : 0;
  static constexpr int X_SIZE = X_INTERP > X_PLANES ? X_INTERP : X_PLANES;
  static constexpr int X_OFF = 6 * ND;
  static constexpr int M_OFF = X_OFF + X_SIZE;
  static constexpr int SMEM = (M_OFF + PL + LXD * LX) * 8;
  static_assert(3 * (S1 + S2) <= 6 * ND, "projection scratch must fit in the field arrays");
};

// (J x J x J) of one element: src (LX^3, global) -> dst (LXD^3; shared or global), scratch X in shared
// memory.  Js = J in shared memory, Jc = J in the constant bank.  All threads of the CTA must call it.
template <int LX, int LXD, int NTHR>
__device__ __forceinline__ void interp_to_fine(const double* __restrict__ src, double* dst, double* X,
                                               const double* Js, const double* Jc, int tid) {
  constexpr int N = LX * LX * LX, PL = LXD * LXD, S1 = LXD * LX * LX, S2 = LXD * LXD * LX;
  double* s0 = X;
  double* s1 = X + N;
  double* s2 = X + N + S1;
  for (int idx = tid; idx < N; idx += NTHR) s0[idx] = __ldg(src + idx);
  __syncthreads();
  for (int idx = tid; idx < S1; idx += NTHR) {     // s1(a,m,n) = sum_l J(a,l) s0(l,m,n)
    const int a = idx % LXD, mn = idx / LXD;
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < LX; l++) s = fma(Js[a + LXD * l], s0[l + LX * mn], s);
    s1[idx] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < S2; idx += NTHR) {     // s2(a,b,n) = sum_m s1(a,m,n) J(b,m)
    const int a = idx % LXD, b = (idx / LXD) % LXD, n = idx / PL;
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < LX; m++) s = fma(s1[a + LXD * (m + LX * n)], Js[b + LXD * m], s);
    s2[idx] = s;
  }
  __syncthreads();
  if (tid < PL) {                                  // dst(a,b,c) = sum_n s2(a,b,n) J(c,n): own column
    double r[LX];
#pragma unroll
    for (int n = 0; n < LX; n++) r[n] = s2[tid + PL * n];
#pragma unroll
    for (int c = 0; c < LXD; c++) {
      double s = 0.0;
#pragma unroll
      for (int n = 0; n < LX; n++) s = fma(r[n], Jc[c + LXD * n], s);
      dst[tid + PL * c] = s;
    }
  }
  // the caller's next barrier (or the next call's first one) orders s2 reads before it is rewritten
}

template <int V>
struct IntC { static constexpr int value = V; };

// (J x J x J) of NB fields of one element at once: src[f] (LX^3, global) -> T + f*ND (LXD^3, shared).  Every
// stage is a set of independent pencil tasks -- a thread pulls the LX (or LXD) values of a pencil into
// registers and produces all outputs of that pencil with J as constant-bank operands, so the only shared
// traffic is the data itself, and one barrier separates the stages for all NB fields together:
//   1. rows (m,n):      s1(a,m,n) = sum_l J(a,l) u(l,m,n)           global -> X (NB * LXD*LX*LX doubles)
//   2. pencils (a,n):   s2(a,b,n) = sum_m s1(a,m,n) J(b,m)          X -> first LX planes of T_f
//   3. columns (a,b):   T(a,b,c)  = sum_n s2(a,b,n) J(c,n)          in place (a column is thread-private)
template <int LX, int LXD, int NTHR, int NB>
__device__ __forceinline__ void interp_batch(const double* const (&src)[NB], double* T, double* X,
                                             const double* Jc, int tid) {
  constexpr int ND = LXD * LXD * LXD, PL = LXD * LXD, S1 = LXD * LX * LX;
  for (int t = tid; t < NB * LX * LX; t += NTHR) {
    const int f = t / (LX * LX), mn = t - f * (LX * LX);
    const double* sp = src[0];               // (no run-time indexing of the pointer array: it would go to local memory)
#pragma unroll
    for (int q = 1; q < NB; q++) sp = (f == q) ? src[q] : sp;
    double u[LX];
#pragma unroll
    for (int l = 0; l < LX; l++) u[l] = __ldg(sp + l + LX * mn);
    double* o = X + f * S1 + LXD * mn;
#pragma unroll
    for (int a = 0; a < LXD; a++) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < LX; l++) s = fma(Jc[a + LXD * l], u[l], s);
      o[a] = s;
    }
  }
  __syncthreads();
  for (int t = tid; t < NB * LXD * LX; t += NTHR) {
    const int a = t % LXD, n = (t / LXD) % LX, f = t / (LXD * LX);
    const double* in = X + f * S1 + a + LXD * LX * n;
    double v[LX];
#pragma unroll
    for (int m = 0; m < LX; m++) v[m] = in[LXD * m];
    double* o = T + f * ND + a + PL * n;
#pragma unroll
    for (int b = 0; b < LXD; b++) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < LX; m++) s = fma(v[m], Jc[b + LXD * m], s);
      o[LXD * b] = s;
    }
  }
  __syncthreads();
  for (int t = tid; t < NB * PL; t += NTHR) {
    const int f = t / PL, ab = t - f * PL;
    double* col = T + f * ND + ab;
    double r[LX];
#pragma unroll
    for (int n = 0; n < LX; n++) r[n] = col[PL * n];
#pragma unroll
    for (int c = 0; c < LXD; c++) {
      double s = 0.0;
#pragma unroll
      for (int n = 0; n < LX; n++) s = fma(r[n], Jc[c + LXD * n], s);
      col[PL * c] = s;
    }
  }
  __syncthreads();
}

template <int LX, int LXD, int MODE, int MAXREG>
__global__ void __launch_bounds__(AdvCfg<LX, LXD, MODE>::NTHR) __maxnreg__(MAXREG)
advop_kernel(const __grid_constant__ AdvParams<LX, LXD> p) {
  using C = AdvCfg<LX, LXD, MODE>;
  constexpr int N = C::N, ND = C::ND, PL = C::PL, NTHR = C::NTHR, S1 = C::S1, S2 = C::S2;
  extern __shared__ __align__(16) double sm[];
  double* T = sm;                 // T[q*ND + .]: q = 0..2 v, 3..5 U_b on the fine grid
  double* X = sm + C::X_OFF;      // interpolation scratch / plane buffers
  double* Ds = sm + C::M_OFF;     // D(i,m) at Ds[i + LXD*m]
  double* Js = Ds + PL;           // J(a,l) at Js[a + LXD*l]

  const int tid = threadIdx.x;
  for (int idx = tid; idx < PL; idx += NTHR) Ds[idx] = p.D[idx];
  if constexpr (C::INTERP)
    for (int idx = tid; idx < LXD * LX; idx += NTHR) Js[idx] = p.J[idx];
  const bool act = tid < PL;
  const int i = act ? tid % LXD : 0, j = act ? tid / LXD : 0;
  const double wij = p.wd[i] * p.wd[j];
  const unsigned flags = p.flags;
  __syncthreads();

  for (int it = blockIdx.x; it < p.nelem; it += gridDim.x) {
    const int e = p.elem_list ? __ldg(p.elem_list + it) : p.elem_base + it;
    const size_t eb = (size_t)e * N, ebd = (size_t)e * ND;

    // ---- phase 0: the six fields on the fine grid ----------------------------------------------------
    if constexpr (C::INTERP) {
      {
        const double* const sv[3] = {p.v[0] + eb, p.v[1] + eb, p.v[2] + eb};
        interp_batch<LX, LXD, NTHR, 3>(sv, T, X, p.J, tid);
      }
      {
        const double* const sb[3] = {p.vb[0] + eb, p.vb[1] + eb, p.vb[2] + eb};
        interp_batch<LX, LXD, NTHR, 3>(sb, T + 3 * ND, X, p.J, tid);
      }
    } else {
#pragma unroll 1
      for (int qf = 0; qf < 6; qf++) {
        const double* src = (qf < 3 ? p.v[qf] : p.vb[qf - 3]) + eb;
        for (int idx = tid; idx < N; idx += NTHR) T[qf * ND + idx] = __ldg(src + idx);
      }
      __syncthreads();
    }

    // ---- phase 1: the k-column of thread (i,j) ----------------------------------------------------------
    double acc[3][LXD];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < LXD; k++) acc[c][k] = 0.0;

#pragma unroll
    for (int k = 0; k < LXD; k++) {
      const int pidx = tid + PL * k;
      double* PB = X + (k & 1) * 6 * PL;     // plane buffers of this plane (ADJOINT only)
      if (act) {
        double g[9];
#pragma unroll
        for (int a = 0; a < 9; a++) g[a] = __ldg(p.G[a] + ebd + pidx);
        const double w3 = wij * p.wd[k];
        double tv[3], tb[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { tv[c] = T[c * ND + pidx]; tb[c] = T[(3 + c) * ND + pidx]; }

        // reference-space derivatives; q = 0..2: U_b components, 3..5: v components (LINEAR only)
        constexpr int NDF = (MODE == ADV_LINEAR) ? 6 : 3;
        double dr[NDF], ds[NDF], dt[NDF];
#pragma unroll
        for (int qd = 0; qd < NDF; qd++) { dr[qd] = 0.0; ds[qd] = 0.0; dt[qd] = 0.0; }
#pragma unroll
        for (int m = 0; m < LXD; m++) {
          const double di = Ds[i + LXD * m], dj = Ds[j + LXD * m];   // one load of D(i,m), D(j,m) for all fields
#pragma unroll
          for (int qd = 0; qd < NDF; qd++) {
            const double* U = T + ((qd < 3) ? (3 + qd) : (qd - 3)) * ND;
            dr[qd] = fma(di, U[m + LXD * j + PL * k], dr[qd]);
            ds[qd] = fma(dj, U[i + LXD * m + PL * k], ds[qd]);
            dt[qd] = fma(p.D[k + LXD * m], U[tid + PL * m], dt[qd]);
          }
        }

        if constexpr (MODE == ADV_ADJOINT) {
          double R0 = 0.0, R1 = 0.0, R2 = 0.0;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double dx = w3 * (g[0] * dr[c] + g[1] * ds[c] + g[2] * dt[c]);
            const double dy = w3 * (g[3] * dr[c] + g[4] * ds[c] + g[5] * dt[c]);
            const double dz = w3 * (g[6] * dr[c] + g[7] * ds[c] + g[8] * dt[c]);
            R0 = fma(tv[c], dx, R0);
            R1 = fma(tv[c], dy, R1);
            R2 = fma(tv[c], dz, R2);
          }
          acc[0][k] += R0; acc[1][k] += R1; acc[2][k] += R2;
          const double cr = w3 * (tb[0] * g[0] + tb[1] * g[3] + tb[2] * g[6]);
          const double cs = w3 * (tb[0] * g[1] + tb[1] * g[4] + tb[2] * g[7]);
          const double ct = w3 * (tb[0] * g[2] + tb[1] * g[5] + tb[2] * g[8]);
#pragma unroll
          for (int c = 0; c < 3; c++) {
            PB[c * PL + tid] = tv[c] * cr;
            PB[(3 + c) * PL + tid] = tv[c] * cs;
            const double ft = tv[c] * ct;
#pragma unroll
            for (int m = 0; m < LXD; m++) acc[c][m] = fma(p.D[k + LXD * m], ft, acc[c][m]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; c++) {
            // u' . grad U_b,c  (pass 0)  +  U_b . grad u'_c  (pass 1)
            const double bx = w3 * (g[0] * dr[c] + g[1] * ds[c] + g[2] * dt[c]);
            const double by = w3 * (g[3] * dr[c] + g[4] * ds[c] + g[5] * dt[c]);
            const double bz = w3 * (g[6] * dr[c] + g[7] * ds[c] + g[8] * dt[c]);
            const double vx = w3 * (g[0] * dr[3 + c] + g[1] * ds[3 + c] + g[2] * dt[3 + c]);
            const double vy = w3 * (g[3] * dr[3 + c] + g[4] * ds[3 + c] + g[5] * dt[3 + c]);
            const double vz = w3 * (g[6] * dr[3 + c] + g[7] * ds[3 + c] + g[8] * dt[3 + c]);
            acc[c][k] = (tv[0] * bx + tv[1] * by + tv[2] * bz) + (tb[0] * vx + tb[1] * vy + tb[2] * vz);
          }
        }
      }
      if constexpr (MODE == ADV_ADJOINT) {
        __syncthreads();
        if (act) {
          double rr[3] = {0.0, 0.0, 0.0}, ss[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int m = 0; m < LXD; m++) {
            const double di = Ds[m + LXD * i], dj = Ds[m + LXD * j];
#pragma unroll
            for (int c = 0; c < 3; c++) {
              rr[c] = fma(di, PB[c * PL + m + LXD * j], rr[c]);
              ss[c] = fma(dj, PB[(3 + c) * PL + i + LXD * m], ss[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; c++) acc[c][k] += rr[c] + ss[c];
        }
      }
    }

    // ---- phase 2: back to the GLL grid + epilogue ----------------------------------------------------------
    auto epilogue = [&](int pt, double o0, double o1, double o2) {
      const size_t gi = eb + pt;
      if (flags & FLAG_ACCUM) {
        p.f[0][gi] -= o0; p.f[1][gi] -= o1; p.f[2][gi] -= o2;
        return;
      }
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if (flags & (FLAG_SOURCES | FLAG_FSTATIC | FLAG_SENS)) {
        const double pv0 = __ldg(p.v[0] + gi), pv1 = __ldg(p.v[1] + gi), pv2 = __ldg(p.v[2] + gi);
        const double b0 = __ldg(p.vb[0] + gi), b1 = __ldg(p.vb[1] + gi), b2 = __ldg(p.vb[2] + gi);
        const double bm = (flags & (FLAG_SOURCES | FLAG_FSTATIC)) ? __ldg(p.B + gi) : 0.0;
        if (flags & FLAG_SOURCES) {
          double ch = __ldg(p.rho + gi);
          if (flags & FLAG_RAMP) {
            if (flags & FLAG_CONVEX_UP) ch = p.f_min + (p.f_max - p.f_min) * ch * (1.0 + p.q) / (ch + p.q);
            else ch = p.f_min + (p.f_max - p.f_min) * ch / (1.0 + p.q * (1.0 - ch));
          }
          f0 = 0.0 - pv0 * ch; f1 = 0.0 - pv1 * ch; f2 = 0.0 - pv2 * ch;
          if (flags & FLAG_FSTATIC) { f0 += __ldg(p.fs[0] + gi); f1 += __ldg(p.fs[1] + gi); f2 += __ldg(p.fs[2] + gi); }
          if (flags & FLAG_LUBE) {
            const double ck = ch * p.K_lube;
            f0 += b0 * ck; f1 += b1 * ck; f2 += b2 * ck;
          }
          f0 *= bm; f1 *= bm; f2 *= bm;
          if (flags & FLAG_CHI_OUT) p.chi_out[gi] = ch;
        } else if (flags & FLAG_FSTATIC) {
          f0 = __ldg(p.fs[0] + gi) * bm; f1 = __ldg(p.fs[1] + gi) * bm; f2 = __ldg(p.fs[2] + gi) * bm;
        }
        if (flags & FLAG_SENS) {
          double sv = b0 * pv0;
          sv = fma(b1, pv1, sv);
          sv = fma(b2, pv2, sv);
          sv = -sv;
          double l = b0 * b0;
          l = fma(b1, b1, l);
          l = fma(b2, b2, l);
          p.sens[gi] = fma(p.K_sens, l, sv);
        }
      }
      p.f[0][gi] = f0 - o0; p.f[1][gi] = f1 - o1; p.f[2][gi] = f2 - o2;
    };

    if constexpr (C::INTERP) {
      __syncthreads();                         // every thread is done with T and the plane buffers
      double* P1 = T;                          // P1_c(i,j,n) = sum_k J(k,n) R_c(i,j,k)
      double* P2 = T + 3 * S2;                 // P2_c(i,m,n) = sum_j J(j,m) P1_c(i,j,n)
      if (act) {
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int n = 0; n < LX; n++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < LXD; k++) s = fma(p.J[k + LXD * n], acc[c][k], s);
            P1[c * S2 + tid + PL * n] = s;
          }
      }
      __syncthreads();
      for (int t = tid; t < 3 * LXD * LX; t += NTHR) {     // pencils (c, i, n): contract j
        const int ii = t % LXD, n = (t / LXD) % LX, c = t / (LXD * LX);
        const double* in = P1 + c * S2 + ii + PL * n;
        double v[LXD];
#pragma unroll
        for (int jj = 0; jj < LXD; jj++) v[jj] = in[LXD * jj];
        double* o = P2 + c * S1 + ii + LXD * LX * n;
#pragma unroll
        for (int m = 0; m < LX; m++) {
          double sacc = 0.0;
#pragma unroll
          for (int jj = 0; jj < LXD; jj++) sacc = fma(p.J[jj + LXD * m], v[jj], sacc);
          o[LXD * m] = sacc;
        }
      }
      __syncthreads();
      // rows (m,n), two tasks per row (first / second half of l): contract i, then the epilogue
      constexpr int LH = (LX + 1) / 2;
      auto row_task = [&](int mn, auto half_c) {
        constexpr int HALF = decltype(half_c)::value;
        constexpr int L0 = HALF * LH, L1 = HALF ? LX : LH;
        double v[3][LXD];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int ii = 0; ii < LXD; ii++) v[c][ii] = P2[c * S1 + ii + LXD * mn];
#pragma unroll
        for (int l = L0; l < L1; l++) {
          double o[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double sacc = 0.0;
#pragma unroll
            for (int ii = 0; ii < LXD; ii++) sacc = fma(p.J[ii + LXD * l], v[c][ii], sacc);
            o[c] = sacc;
          }
          epilogue(l + LX * mn, o[0], o[1], o[2]);
        }
      };
      for (int t = tid; t < 2 * LX * LX; t += NTHR) {
        if (t & 1) row_task(t >> 1, IntC<1>{});
        else row_task(t >> 1, IntC<0>{});
      }
      __syncthreads();                         // T is rewritten by the next element
    } else {
      if (act) {
#pragma unroll
        for (int k = 0; k < LXD; k++) epilogue(tid + PL * k, acc[0][k], acc[1][k], acc[2][k]);
      }
      __syncthreads();
    }
  }
}

// coef_GL of init_dealias (adv_adjoint_dealias.f90:153-161): the nine GLL geometric factors interpolated to
// the fine grid, element by element; grid-stride over (field, element) pairs.
template <int LX, int LXD>
struct GeomInterpParams {
  double J[LXD * LX];
  const double* src[9];
  double* dst[9];
  int nelv;
};

template <int LX, int LXD>
__global__ void __launch_bounds__(((LXD * LXD + 31) / 32) * 32)
geom_to_fine_kernel(const __grid_constant__ GeomInterpParams<LX, LXD> p) {
  constexpr int N = LX * LX * LX, ND = LXD * LXD * LXD, PL = LXD * LXD;
  constexpr int NTHR = ((PL + 31) / 32) * 32;
  extern __shared__ __align__(16) double sm[];
  double* X = sm;
  double* Js = sm + N + LXD * LX * LX + LXD * LXD * LX;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < LXD * LX; idx += NTHR) Js[idx] = p.J[idx];
  __syncthreads();
  const long long total = 9ll * p.nelv;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int a = (int)(w % 9);
    const long long e = w / 9;
    interp_to_fine<LX, LXD, NTHR>(p.src[a] + e * N, p.dst[a] + e * ND, X, Js, p.J, tid);
    __syncthreads();
  }
}

}  // namespace b200
