// Dealiased adjoint advection operator for lx = 8 / lxd = 12 on the FP64 tensor cores (sm_100a): the same operator
// as advop_kernel<8, 12, ADV_ADJOINT> (reference adjoint/adv_adjoint_dealias.f90:235-462), with
// every contraction written as a batched small GEMM on `mma.sync.m8n8k4.f64`.
//
// Why: the column-per-thread kernel is latency-bound at 2.5 warps per scheduler (168 registers x 160 threads, 104 KB
// per element; ncu r02h: issue 25 %, fp64 pipe 26 %, 45 % of the shared-memory wavefront peak) because every DFMA of
// an r/s contraction needs its own shared-memory operand.  A DMMA takes one 64-bit fragment load per 256 FMAs, needs
// ~60 registers per warp instead of 168, and so 16 warps share one SM with the whole element (15 fine-grid arrays)
// in shared memory.
//
// Formulation.  With J (12x8) the GLL -> Gauss-Legendre interpolation, D (12x12) the fine-grid derivative and
// DJ = D J (12x8, formed on the host), per element:
//   forward   T_q   = (J  x J  x J ) q            q = v_1..3, U_1..3          GLL_to_GL%map      (:360-367)
//             d_r U = (DJ x J  x J ) U, d_s U = (J x DJ x J) U, d_t U = (J x J x DJ) U   == opgrad's D applied to T_U
//   point     R_d   = sum_c v_c w3 (G_rd d_r + G_sd d_s + G_td d_t) U_c        opgrad + vdot3     (:372-381)
//             Fr_c = v_c c_r, Fs_c = v_c c_s, Ft_c = v_c c_t,  c_r = w3 sum_k U_k G_rk   (the 9 cdtp arguments, :394-453)
//   backward  out_c = (J^T x J^T x J^T) R_c + (DJ^T x J^T x J^T) Fr_c + (J^T x DJ^T x J^T) Fs_c + (J^T x J^T x DJ^T) Ft_c
//             == map(., Xh_GLL) of R_c + cdtp(...)  (:384-392; D^T then J^T along one axis = (DJ)^T)
// Each product is evaluated axis by axis, IN PLACE in the arrays (a warp tile reads all the pencils of its 8 batch
// entries before it stores them), sums that share the remaining axes are accumulated in the DMMA accumulators:
//   F1 (r):  A1 = J_r q,  A2 = DJ_r U                      F2 (s):  B1 = J_s A1, B2 = DJ_s A1, B3 = J_s A2
//   F3 (t):  T = J_t B1, d_t = DJ_t B1, d_s = J_t B2, d_r = J_t B3
//   T1 (t):  X1 = J^T_t R + DJ^T_t Ft,  X2 = J^T_t Fr,  X3 = J^T_t Fs
//   T2 (s):  Y1 = J^T_s X1 + DJ^T_s X3, Y2 = J^T_s X2      T3 (r):  out = J^T_r Y1 + DJ^T_r Y2
// 3060 DMMAs per element (75 % useful forward: 12 outputs in two 8-row tiles; 100 % backward) against 750 k DFMAs.
//
// Fragments (PTX m8n8k4, lane = 4 g + q): A(row g, col q), B(row q, col g), C(row g, cols 2q, 2q+1).  The DATA
// fragment of a tile is always "contraction index 4 ks + q of batch entry g", whichever operand it is; the MATRIX
// fragment is always M(output g, contraction q).  Matrix as A: C = (output g; batch 2q, 2q+1); matrix as B:
// C = (batch g; output 2q, 2q+1) -- chosen per stage so that the two C values are adjacent in memory (128-bit stores).
//
// Shared memory: 15 arrays [k][j][i] with row stride 12 and plane stride 148 doubles (148 = 4 mod 16: the four
// contraction indices of a t-stage fragment load fall into different banks), 216.8 KB; one CTA of 512 threads per SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "adjrhs_common.cuh"   // FLAG_*

namespace b200 {

struct AdvMmaParams {
  double J[96];            // J(a,l) at J[a + 12*l]
  double DJ[96];           // (D J)(a,l) at DJ[a + 12*l]
  double wd[12];           // fine-grid quadrature weights
  const double* v[3];      // adjoint velocity (GLL)
  const double* vb[3];     // base flow (GLL)
  const double* G[9];      // geometric factors on the fine grid, nelv*1728 each
  double* f[3];
  const double* rho;
  const double* B;
  const double* fs[3];
  double* sens;
  double* chi_out;
  const int* elem_list;
  int nelem;
  int elem_base;
  unsigned flags;
  double f_min, f_max, q, K_lube, K_sens;
};

struct AdvMmaCfg {
  static constexpr int LX = 8, LXD = 12, N = 512, ND = 1728, PL = 144;
  static constexpr int PS = 148;                  // plane stride in shared memory
  static constexpr int AS = LXD * PS;             // array stride
  static constexpr int NARR = 15;                 // TV 0..2, TB 3..5, DR 6..8, DS 9..11, DT 12..14
  static constexpr int NFRAG = 14;                // J: 0..3 (mt*2+ks), DJ: 4..7, J^T: 8..10 (ks), DJ^T: 11..13
  static constexpr int FT_OFF = NARR * AS;
  static constexpr int W_OFF = FT_OFF + NFRAG * 32;
  static constexpr int SMEM = (W_OFF + 12) * 8;
  static constexpr int NTHR = 512, NWARP = 16;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st2(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
__device__ __forceinline__ void adv_l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(AdvMmaCfg::NTHR, 1)
advop_mma_kernel(const __grid_constant__ AdvMmaParams p) {
  using C = AdvMmaCfg;
  constexpr int PS = C::PS, AS = C::AS, N = C::N, ND = C::ND, PL = C::PL;
  extern __shared__ __align__(16) double sm[];
  double* FT = sm + C::FT_OFF;
  double* W = sm + C::W_OFF;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const unsigned flags = p.flags;

  // matrix fragments [fragment][lane] and weights
  for (int idx = tid; idx < C::NFRAG * 32; idx += C::NTHR) {
    const int fid = idx >> 5, gg = (idx & 31) >> 2, qq = idx & 3;
    double val;
    if (fid < 8) {                     // forward: M(output a = 8 mt + g, contraction l = 4 ks + q), rows >= 12 are zero
      const int mt = (fid & 3) >> 1, ks = fid & 1, o = 8 * mt + gg, c = 4 * ks + qq;
      val = (o < 12) ? (fid < 4 ? p.J[o + 12 * c] : p.DJ[o + 12 * c]) : 0.0;
    } else {                           // backward: M(output l = g, contraction a = 4 ks + q) = J(a, l)
      const int ks = (fid - 8) % 3, c = 4 * ks + qq;
      val = (fid < 11) ? p.J[c + 12 * gg] : p.DJ[c + 12 * gg];
    }
    FT[idx] = val;
  }
  if (tid < 12) W[tid] = p.wd[tid];

  // GLL point of this thread (load of the six fields, epilogue)
  const int gl = tid & 7, gm = (tid >> 3) & 7, gn = tid >> 6;
  const int goff = gl + 12 * gm + PS * gn;

  auto elem_of = [&](int it) { return p.elem_list ? __ldg(p.elem_list + it) : p.elem_base + it; };
  double un[6];
  if ((int)blockIdx.x < p.nelem) {
    const size_t eb = (size_t)elem_of(blockIdx.x) * N + tid;
#pragma unroll
    for (int c = 0; c < 3; c++) { un[c] = __ldg(p.v[c] + eb); un[3 + c] = __ldg(p.vb[c] + eb); }
  }
  __syncthreads();

  for (int it = blockIdx.x; it < p.nelem; it += gridDim.x) {
    const int e = elem_of(it);
    const size_t eb = (size_t)e * N, ebd = (size_t)e * ND;
#pragma unroll
    for (int f = 0; f < 6; f++) sm[f * AS + goff] = un[f];
    if (tid < 9) adv_l2_prefetch(p.G[tid] + ebd, ND * 8);
    __syncthreads();

    // ---- F1: r axis.  batch (m, n): tile t = n, entry g = m; output a = 2q, 2q+1 (+8 nt): matrix as B ------------
    for (int tk = warp; tk < 48; tk += C::NWARP) {
      const int fo = tk >> 3, t = tk & 7;
      const bool isb = fo < 3;                       // base-flow fields first (twice the work)
      double* arr = sm + (isb ? 3 + fo : fo - 3) * AS + 12 * g + PS * t;
      const double d0 = arr[q], d1 = arr[4 + q];
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, d0, FT[(nt * 2 + 0) * 32 + lane]);
        dmma884(c0, c1, d1, FT[(nt * 2 + 1) * 32 + lane]);
        if (nt == 0 || q < 2) st2(arr + 8 * nt + 2 * q, c0, c1);
      }
      if (isb) {
        double* dr = sm + (6 + fo) * AS + 12 * g + PS * t;
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          double c0 = 0.0, c1 = 0.0;
          dmma884(c0, c1, d0, FT[(4 + nt * 2 + 0) * 32 + lane]);
          dmma884(c0, c1, d1, FT[(4 + nt * 2 + 1) * 32 + lane]);
          if (nt == 0 || q < 2) st2(dr + 8 * nt + 2 * q, c0, c1);
        }
      }
    }
    __syncthreads();

    // ---- F2: s axis.  batch beta = a + 12 n (96); output b = g (+8 mt): matrix as A ----------------------------------
    for (int tk = warp; tk < 72; tk += C::NWARP) {
      const int fo = tk / 12, t = tk - 12 * fo;
      const bool isb = fo < 3;
      const int bl = 8 * t + g, bs = 8 * t + 2 * q;                 // load / store batch entry
      const int lo = (bl % 12) + PS * (bl / 12) + 12 * q;           // + 48 per k-step
      const int so = (bs % 12) + PS * (bs / 12) + 12 * g;           // + 96 for mt = 1
      double* arr = sm + (isb ? 3 + fo : fo - 3) * AS;
      const double d0 = arr[lo], d1 = arr[lo + 48];
      double e0 = 0.0, e1 = 0.0;
      double* dr = sm + (6 + fo) * AS;
      double* ds = sm + (9 + fo) * AS;
      if (isb) { e0 = dr[lo]; e1 = dr[lo + 48]; }
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        const double j0 = FT[(mt * 2 + 0) * 32 + lane], j1 = FT[(mt * 2 + 1) * 32 + lane];
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, j0, d0);
        dmma884(c0, c1, j1, d1);
        if (mt == 0 || g < 4) st2(arr + so + 96 * mt, c0, c1);
        if (isb) {
          double b0 = 0.0, b1 = 0.0, a0 = 0.0, a1 = 0.0;
          dmma884(b0, b1, FT[(4 + mt * 2 + 0) * 32 + lane], d0);
          dmma884(b0, b1, FT[(4 + mt * 2 + 1) * 32 + lane], d1);
          dmma884(a0, a1, j0, e0);
          dmma884(a0, a1, j1, e1);
          if (mt == 0 || g < 4) { st2(ds + so + 96 * mt, b0, b1); st2(dr + so + 96 * mt, a0, a1); }
        }
      }
    }
    __syncthreads();

    // ---- F3: t axis.  batch beta = a + 12 b (144, contiguous); output c = g (+8 mt): matrix as A ------------------
    for (int tk = warp; tk < 108; tk += C::NWARP) {
      const int fo = tk / 18, t = tk - 18 * fo;
      const bool isb = fo < 3;
      const int lo = 8 * t + g + PS * q;                            // + 4 PS per k-step
      const int so = 8 * t + 2 * q + PS * g;                        // + 8 PS for mt = 1
      double* arr = sm + (isb ? 3 + fo : fo - 3) * AS;
      double* dr = sm + (6 + fo) * AS;
      double* ds = sm + (9 + fo) * AS;
      double* dt = sm + (12 + fo) * AS;
      const double d0 = arr[lo], d1 = arr[lo + 4 * PS];
      double r0 = 0.0, r1 = 0.0, s0 = 0.0, s1 = 0.0;
      if (isb) { r0 = dr[lo]; r1 = dr[lo + 4 * PS]; s0 = ds[lo]; s1 = ds[lo + 4 * PS]; }
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        const double j0 = FT[(mt * 2 + 0) * 32 + lane], j1 = FT[(mt * 2 + 1) * 32 + lane];
        const bool on = (mt == 0 || g < 4);
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, j0, d0);
        dmma884(c0, c1, j1, d1);
        if (on) st2(arr + so + 8 * PS * mt, c0, c1);
        if (isb) {
          double t0 = 0.0, t1 = 0.0, a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
          dmma884(t0, t1, FT[(4 + mt * 2 + 0) * 32 + lane], d0);
          dmma884(t0, t1, FT[(4 + mt * 2 + 1) * 32 + lane], d1);
          dmma884(a0, a1, j0, s0);
          dmma884(a0, a1, j1, s1);
          dmma884(b0, b1, j0, r0);
          dmma884(b0, b1, j1, r1);
          if (on) {
            st2(dt + so + 8 * PS * mt, t0, t1);
            st2(ds + so + 8 * PS * mt, a0, a1);
            st2(dr + so + 8 * PS * mt, b0, b1);
          }
        }
      }
    }
    __syncthreads();

    // ---- point-wise stage on the fine grid (in place: R -> TV, Fr -> DR, Fs -> DS, Ft -> DT) -------------------------
#pragma unroll 2
    for (int pt = tid; pt < ND; pt += C::NTHR) {
      const int c = pt / PL, ab = pt - PL * c, b = ab / 12, a = ab - 12 * b;
      double* P = sm + ab + PS * c;
      double gg[9];
#pragma unroll
      for (int x = 0; x < 9; x++) gg[x] = __ldg(p.G[x] + ebd + pt);
      const double w3 = W[a] * W[b] * W[c];
      double tv[3], tb[3];
#pragma unroll
      for (int x = 0; x < 3; x++) { tv[x] = P[x * AS]; tb[x] = P[(3 + x) * AS]; }
      double R0 = 0.0, R1 = 0.0, R2 = 0.0;
#pragma unroll
      for (int x = 0; x < 3; x++) {
        const double dr = P[(6 + x) * AS], ds = P[(9 + x) * AS], dt = P[(12 + x) * AS];
        const double dx = w3 * (gg[0] * dr + gg[1] * ds + gg[2] * dt);
        const double dy = w3 * (gg[3] * dr + gg[4] * ds + gg[5] * dt);
        const double dz = w3 * (gg[6] * dr + gg[7] * ds + gg[8] * dt);
        R0 = fma(tv[x], dx, R0);
        R1 = fma(tv[x], dy, R1);
        R2 = fma(tv[x], dz, R2);
      }
      const double cr = w3 * (tb[0] * gg[0] + tb[1] * gg[3] + tb[2] * gg[6]);
      const double cs = w3 * (tb[0] * gg[1] + tb[1] * gg[4] + tb[2] * gg[7]);
      const double ct = w3 * (tb[0] * gg[2] + tb[1] * gg[5] + tb[2] * gg[8]);
      P[0 * AS] = R0; P[1 * AS] = R1; P[2 * AS] = R2;
#pragma unroll
      for (int x = 0; x < 3; x++) {
        P[(6 + x) * AS] = tv[x] * cr;
        P[(9 + x) * AS] = tv[x] * cs;
        P[(12 + x) * AS] = tv[x] * ct;
      }
    }
    __syncthreads();

    // ---- T1: t axis, K = 12.  batch beta = i + 12 j; output n = g: matrix as A.  part 0: X1 -> R, part 1: X2, X3 -----
    for (int tk = warp; tk < 108; tk += C::NWARP) {
      const int c = tk / 36, r = tk - 36 * c, t = r >> 1, part = r & 1;
      const int lo = 8 * t + g + PS * q;
      const int so = 8 * t + 2 * q + PS * g;
      double* a0 = sm + (part ? 6 + c : c) * AS;           // Fr | R
      double* a1 = sm + (part ? 9 + c : 12 + c) * AS;      // Fs | Ft
      double x[3], y[3];
#pragma unroll
      for (int ks = 0; ks < 3; ks++) { x[ks] = a0[lo + 4 * PS * ks]; y[ks] = a1[lo + 4 * PS * ks]; }
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      if (part == 0) {
#pragma unroll
        for (int ks = 0; ks < 3; ks++) {
          dmma884(c0, c1, FT[(8 + ks) * 32 + lane], x[ks]);
          dmma884(c0, c1, FT[(11 + ks) * 32 + lane], y[ks]);
        }
        st2(a0 + so, c0, c1);
      } else {
#pragma unroll
        for (int ks = 0; ks < 3; ks++) {
          const double jt = FT[(8 + ks) * 32 + lane];
          dmma884(c0, c1, jt, x[ks]);
          dmma884(e0, e1, jt, y[ks]);
        }
        st2(a0 + so, c0, c1);
        st2(a1 + so, e0, e1);
      }
    }
    __syncthreads();

    // ---- T2: s axis, K = 12.  batch beta = i + 12 n (96); output m = g.  part 0: Y1 -> R, part 1: Y2 -> Fr -------------
    for (int tk = warp; tk < 72; tk += C::NWARP) {
      const int c = tk / 24, r = tk - 24 * c, t = r >> 1, part = r & 1;
      const int bl = 8 * t + g, bs = 8 * t + 2 * q;
      const int lo = (bl % 12) + PS * (bl / 12) + 12 * q;            // + 48 per k-step
      const int so = (bs % 12) + PS * (bs / 12) + 12 * g;
      double c0 = 0.0, c1 = 0.0;
      if (part == 0) {
        double* R = sm + c * AS;
        const double* X3 = sm + (9 + c) * AS;
        double x[3], y[3];
#pragma unroll
        for (int ks = 0; ks < 3; ks++) { x[ks] = R[lo + 48 * ks]; y[ks] = X3[lo + 48 * ks]; }
#pragma unroll
        for (int ks = 0; ks < 3; ks++) {
          dmma884(c0, c1, FT[(8 + ks) * 32 + lane], x[ks]);
          dmma884(c0, c1, FT[(11 + ks) * 32 + lane], y[ks]);
        }
        st2(R + so, c0, c1);
      } else {
        double* X2 = sm + (6 + c) * AS;
        double x[3];
#pragma unroll
        for (int ks = 0; ks < 3; ks++) x[ks] = X2[lo + 48 * ks];
#pragma unroll
        for (int ks = 0; ks < 3; ks++) dmma884(c0, c1, FT[(8 + ks) * 32 + lane], x[ks]);
        st2(X2 + so, c0, c1);
      }
    }
    __syncthreads();

    // ---- T3: r axis, K = 12.  batch (m, n): tile t = n, entry g = m; output l = 2q, 2q+1: matrix as B -------------------
    for (int tk = warp; tk < 24; tk += C::NWARP) {
      const int c = tk >> 3, t = tk & 7;
      double* R = sm + c * AS + 12 * g + PS * t;
      const double* Y2 = sm + (6 + c) * AS + 12 * g + PS * t;
      double x[3], y[3];
#pragma unroll
      for (int ks = 0; ks < 3; ks++) { x[ks] = R[4 * ks + q]; y[ks] = Y2[4 * ks + q]; }
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 3; ks++) {
        dmma884(c0, c1, x[ks], FT[(8 + ks) * 32 + lane]);
        dmma884(c0, c1, y[ks], FT[(11 + ks) * 32 + lane]);
      }
      st2(R + 2 * q, c0, c1);
    }
    // the next element's fields travel while this one is finished
    if (it + (int)gridDim.x < p.nelem) {
      const size_t nb = (size_t)elem_of(it + gridDim.x) * N + tid;
#pragma unroll
      for (int c = 0; c < 3; c++) { un[c] = __ldg(p.v[c] + nb); un[3 + c] = __ldg(p.vb[c] + nb); }
    }
    __syncthreads();

    // ---- epilogue at the GLL point of this thread ------------------------------------------------------------------------
    {
      const double o0 = sm[0 * AS + goff], o1 = sm[1 * AS + goff], o2 = sm[2 * AS + goff];
      const size_t gi = eb + tid;
      if (flags & FLAG_ACCUM) {
        p.f[0][gi] -= o0; p.f[1][gi] -= o1; p.f[2][gi] -= o2;
      } else {
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
      }
    }
    __syncthreads();    // the arrays are rewritten by the next element
  }
}

}  // namespace b200
