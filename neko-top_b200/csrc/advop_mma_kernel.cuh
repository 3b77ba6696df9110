// Dealiased adjoint advection operator for lx = 8 / lxd = 12 on the FP64 tensor cores (sm_100a): the same operator
// as advop_kernel<8, 12, ADV_ADJOINT> (reference adjoint/adv_adjoint_dealias.f90:235-462), with
// every contraction written as a batched small GEMM on `mma.sync.m8n8k4.f64`.
//
// Why: the column-per-thread kernel is latency-bound at 2.5 warps per scheduler (168 registers x 160 threads, 104 KB
// per element; ncu r02h: issue 25 %, fp64 pipe 26 %, 45 % of the shared-memory wavefront peak) because every DFMA of
// an r/s contraction needs its own shared-memory operand.  A DMMA takes one 64-bit fragment load per 256 FMAs, needs
// ~60 registers per warp instead of 168, and so 18 warps share one SM with the whole element (15 fine-grid arrays)
// in shared memory.  Measured (profiles/r02D, r02G, r02J, r02K): fused dealiased step 6.55 -> 2.69 ms at 32^3 elements
// (2.56 -> 6.24 GDOF/s), un-fused drop-in 5.77 -> 2.58 ms.
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
//            (A1 of the base flow goes to the d_t array, free until F3, so that F1 splits into 72 equal tasks)
//   F3 (t):  T = J_t B1, d_t = DJ_t B1, d_s = J_t B2, d_r = J_t B3
//   T1 (t):  X1 = J^T_t R + DJ^T_t Ft,  X2 = J^T_t Fr,  X3 = J^T_t Fs
//   T2 (s):  Y1 = J^T_s X1 + DJ^T_s X3, Y2 = J^T_s X2      T3 (r):  out = J^T_r Y1 + DJ^T_r Y2
// 2880 DMMAs per element (forward: 12 outputs in two 8-row tiles, the leftover rows of J and DJ stacked into one tile
// where they multiply the same data; 100 % useful backward) against 750 k DFMAs.
//
// Fragments (PTX m8n8k4, lane = 4 g + q): A(row g, col q), B(row q, col g), C(row g, cols 2q, 2q+1).  The DATA
// fragment of a tile is always "contraction index 4 ks + q of batch entry g", whichever operand it is; the MATRIX
// fragment is always M(output g, contraction q).  Matrix as A: C = (output g; batch 2q, 2q+1); matrix as B:
// C = (batch g; output 2q, 2q+1) -- chosen per stage so that the two C values are adjacent in memory (128-bit stores).
//
// Shared memory: 15 arrays [k][j][i] with row stride 12 and plane stride 148 doubles (148 = 4 mod 16: the four
// contraction indices of a t-stage fragment load fall into different banks), 220 KB with the fragment tables; one CTA
// of 18 warps per SM (96 registers).  Stages per element, separated by CTA barriers:
//   F1 (data fragments straight from global memory, loaded one element ahead) | F2 | F3 -> point-wise -> T1 (one
//   8-column tile per warp: these three only couple the 12 planes of a column, so there is no CTA barrier inside) |
//   T2 | T3 | epilogue (one GLL point per thread).
// The matrix fragments sit in registers for a whole stage (re-loading them per tile was 45 % of the load wavefronts);
// in the small stages a warp loads the fragments of all its tasks before the first DMMA; the epilogue's inputs and the
// next element's F1 fragments are loaded two stages early with volatile loads; the fine-grid geometry of the element
// is prefetched into L2 when the element starts.  ncu (r02J): DMMA pipe 49 % busy, shared-memory wavefronts 52 %; the
// rest is the point-wise phase (21 %, bound by shared-memory / L1 wavefronts: 27 shared + 9 global 8-byte accesses per
// point) and barrier waits (18 tiles and 18 warps over 4 schedulers: two schedulers carry 25 % more tensor work).
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
  static constexpr int NFRAG = 30;                // J: 0..3 (mt*2+ks), DJ: 4..7, J^T: 8..10 (ks), DJ^T: 11..13 with output row
                                                  // pi(g) (matrix as A); 14..27: the same with output row g (matrix as B);
                                                  // 28, 29 (ks): rows pi(g) < 4 -> J row 8 + pi(g), else DJ row 4 + pi(g)
  static constexpr int FT_OFF = NARR * AS;
  static constexpr int W_OFF = FT_OFF + NFRAG * 32;
  static constexpr int SMEM = (W_OFF + 12) * 8;          // 220.4 KB
  static constexpr int NTHR = 576, NWARP = 18;     // 18 warps = the 18 column tiles of the fused F3 / point-wise / T1 stage
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st2(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
// a load the compiler must leave where it is written (issued two stages before its use)
__device__ __forceinline__ double ldg_pinned(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_pinned_rw(const double* p) {      // for data this kernel also writes (f, in/out)
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
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
  const int pg = (g & 4) | ((g & 1) << 1) | ((g >> 1) & 1);     // pi(g)
  const unsigned flags = p.flags;

  // matrix fragments [fragment][lane] and weights.  Fragments 0..13 carry output row pi(g), pi = (0 2 1 3 4 6 5 7):
  // the two g of a quarter warp then store rows two apart (24 or 296 words = 4 mod 8 sixteen-byte banks) and the
  // 128-bit stores of a C fragment are conflict-free; with consecutive rows (12 / 148 words = 6 / 2 mod 8) every one
  // of them was a 2-way conflict (ncu r02D: 43 % of the store wavefronts).
  for (int idx = tid; idx < C::NFRAG * 32; idx += C::NTHR) {
    const int fr = idx >> 5, gg0 = (idx & 31) >> 2, qq = idx & 3;
    const int fid = fr % 14;
    const int gg = (fr < 14 || fr >= 28) ? ((gg0 & 4) | ((gg0 & 1) << 1) | ((gg0 >> 1) & 1)) : gg0;
    double val;
    if (fr >= 28) {                    // the four leftover rows of J and of DJ in one 8-row tile (they multiply the same data)
      const int c = 4 * (fr - 28) + qq;
      val = (gg < 4) ? p.J[8 + gg + 12 * c] : p.DJ[4 + gg + 12 * c];
    } else if (fid < 8) {                     // forward: M(output a = 8 mt + row, contraction l = 4 ks + q), rows >= 12 are zero
      const int mt = (fid & 3) >> 1, ks = fid & 1, o = 8 * mt + gg, c = 4 * ks + qq;
      val = (o < 12) ? (fid < 4 ? p.J[o + 12 * c] : p.DJ[o + 12 * c]) : 0.0;
    } else {                           // backward: M(output l = row, contraction a = 4 ks + q) = J(a, l)
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
  const bool gth = tid < N;          // threads that own a GLL point
  // F1's data fragments come straight from global memory, loaded one element ahead: task j of this warp (tk = warp +
  // 18 j -> kind, field fo, plane t) needs l = q, 4 + q of row m = pi(g) of plane t
  double uf0[4], uf1[4];
  auto load_f1_fragments = [&](size_t ebase) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int tk = warp + C::NWARP * j, kind = tk / 24, r = tk - 24 * kind, fo = r >> 3, t = r & 7;
      const double* src = (kind < 2 ? p.vb[fo] : p.v[fo]) + ebase + 8 * pg + 64 * t + q;
      uf0[j] = ldg_pinned(src); uf1[j] = ldg_pinned(src + 4);
    }
  };
#pragma unroll
  for (int j = 0; j < 4; j++) { uf0[j] = 0.0; uf1[j] = 0.0; }
  if ((int)blockIdx.x < p.nelem) load_f1_fragments((size_t)elem_of(blockIdx.x) * N);
  __syncthreads();

  for (int it = blockIdx.x; it < p.nelem; it += gridDim.x) {
    const int e = elem_of(it);
    const size_t eb = (size_t)e * N, ebd = (size_t)e * ND;
    double ep_bm = 0.0, ep_rho = 0.0, ep_f[3] = {0.0, 0.0, 0.0};
    if (tid < 9) adv_l2_prefetch(p.G[tid] + ebd, ND * 8);

    // ---- F1: r axis.  batch (m, n): tile t = n, entry g = row m = pi(g); output a = 2q, 2q+1 (+8 nt): matrix as B ---
    // 72 equal tasks (4 per warp): base flow A1 = J_r U -> DT array, A2 = DJ_r U -> DR array, adjoint velocity -> TV;
    // the data fragments were loaded from global memory one element ahead (no staging of the GLL fields, no barrier)
    {
      double mj[4], md[4];
#pragma unroll
      for (int x = 0; x < 4; x++) { mj[x] = FT[(14 + x) * 32 + lane]; md[x] = FT[(18 + x) * 32 + lane]; }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int tk = warp + C::NWARP * j, kind = tk / 24, r = tk - 24 * kind, fo = r >> 3, t = r & 7;
        double* dst = sm + (kind == 0 ? 12 + fo : kind == 1 ? 6 + fo : fo) * AS + 12 * pg + PS * t;
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          double c0 = 0.0, c1 = 0.0;
          dmma884(c0, c1, uf0[j], kind == 1 ? md[nt * 2 + 0] : mj[nt * 2 + 0]);
          dmma884(c0, c1, uf1[j], kind == 1 ? md[nt * 2 + 1] : mj[nt * 2 + 1]);
          if (nt == 0 || q < 2) st2(dst + 8 * nt + 2 * q, c0, c1);
        }
      }
    }
    __syncthreads();

    {
      double mj[4], md[4];     // J, DJ as A operand with output row pi(g): [mt*2 + ks]
#pragma unroll
      for (int x = 0; x < 4; x++) { mj[x] = FT[x * 32 + lane]; md[x] = FT[(4 + x) * 32 + lane]; }
      const double mh[2] = {FT[28 * 32 + lane], FT[29 * 32 + lane]};   // stacked leftover rows of J and DJ

      // ---- F2: s axis.  batch beta = a + 12 n (96); output b = pi(g) (+8 mt): matrix as A ---------------------------
      // per warp 2 base-flow tasks (B1 = J_s A1 -> TB, B2 = DJ_s A1 -> DS, B3 = J_s A2 -> DR in place) and 2 adjoint-velocity
      // tasks (in place); all fragment loads first
      {
        double d0[2], d1[2], e0[2], e1[2], v0[2], v1[2];
        int so[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int tk = warp + C::NWARP * j, fo = tk / 12, t = tk - 12 * fo;
          const int bl = 8 * t + g, bs = 8 * t + 2 * q;               // load / store batch entry
          const int lo = (bl % 12) + PS * (bl / 12) + 12 * q;         // + 48 per k-step
          so[j] = (bs % 12) + PS * (bs / 12) + 12 * pg;               // + 96 for mt = 1
          const double* a1 = sm + (12 + fo) * AS;
          const double* dr = sm + (6 + fo) * AS;
          const double* tv = sm + fo * AS;
          d0[j] = a1[lo]; d1[j] = a1[lo + 48]; e0[j] = dr[lo]; e1[j] = dr[lo + 48];
          v0[j] = tv[lo]; v1[j] = tv[lo + 48];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int tk = warp + C::NWARP * j, fo = tk / 12;
          double* tb = sm + (3 + fo) * AS + so[j];
          double* dr = sm + (6 + fo) * AS + so[j];
          double* ds = sm + (9 + fo) * AS + so[j];
          double* tv = sm + fo * AS + so[j];
          {                                                           // rows 0..7
            double c0 = 0.0, c1 = 0.0, b0 = 0.0, b1 = 0.0, a0 = 0.0, a1v = 0.0, w0 = 0.0, w1 = 0.0;
            dmma884(c0, c1, mj[0], d0[j]);
            dmma884(b0, b1, md[0], d0[j]);
            dmma884(a0, a1v, mj[0], e0[j]);
            dmma884(w0, w1, mj[0], v0[j]);
            dmma884(c0, c1, mj[1], d1[j]);
            dmma884(b0, b1, md[1], d1[j]);
            dmma884(a0, a1v, mj[1], e1[j]);
            dmma884(w0, w1, mj[1], v1[j]);
            st2(tb, c0, c1);
            st2(ds, b0, b1);
            st2(dr, a0, a1v);
            st2(tv, w0, w1);
          }
          {                                                           // rows 8..11: J A1 and DJ A1 share one stacked tile
            double h0 = 0.0, h1 = 0.0, a0 = 0.0, a1v = 0.0, w0 = 0.0, w1 = 0.0;
            dmma884(h0, h1, mh[0], d0[j]);
            dmma884(a0, a1v, mj[2], e0[j]);
            dmma884(w0, w1, mj[2], v0[j]);
            dmma884(h0, h1, mh[1], d1[j]);
            dmma884(a0, a1v, mj[3], e1[j]);
            dmma884(w0, w1, mj[3], v1[j]);
            st2(g < 4 ? tb + 96 : ds + 48, h0, h1);                   // g >= 4: pi(g) = 4 + r -> row 8 + r = pi(g) + 4
            if (g < 4) { st2(dr + 96, a0, a1v); st2(tv + 96, w0, w1); }
          }
        }
      }
      __syncthreads();

      // ---- F3 -> point-wise -> T1 on one tile of 8 (i,j) columns per warp, no CTA barrier in between: the t-axis stages
      // and the point-wise work only couple the 12 planes of a column, so a warp carries its 96 points through all
      // three.  (The warps stay roughly in phase, which is what the FP64 pipe wants: DFMA and DMMA alternate badly.)
      // F3: t axis.  batch beta = a + 12 b (contiguous); output c = pi(g) (+8 mt): matrix as A
      const int t = warp;
      {
        const int lo = 8 * t + g + PS * q;                            // + 4 PS per k-step
        const int so = 8 * t + 2 * q + PS * pg;                       // + 8 PS for mt = 1
#pragma unroll 1
        for (int fo = 0; fo < 3; fo++) {                              // base flow: T, d_t, d_s, d_r
          double* arr = sm + (3 + fo) * AS;
          double* dr = sm + (6 + fo) * AS;
          double* ds = sm + (9 + fo) * AS;
          double* dt = sm + (12 + fo) * AS;
          const double d0 = arr[lo], d1 = arr[lo + 4 * PS];
          const double r0 = dr[lo], r1 = dr[lo + 4 * PS], s0 = ds[lo], s1 = ds[lo + 4 * PS];
          __syncwarp();
          {                                                           // planes 0..7
            double c0 = 0.0, c1 = 0.0, t0 = 0.0, t1 = 0.0, a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            dmma884(c0, c1, mj[0], d0);
            dmma884(t0, t1, md[0], d0);
            dmma884(a0, a1, mj[0], s0);
            dmma884(b0, b1, mj[0], r0);
            dmma884(c0, c1, mj[1], d1);
            dmma884(t0, t1, md[1], d1);
            dmma884(a0, a1, mj[1], s1);
            dmma884(b0, b1, mj[1], r1);
            st2(arr + so, c0, c1);
            st2(dt + so, t0, t1);
            st2(ds + so, a0, a1);
            st2(dr + so, b0, b1);
          }
          {                                                           // planes 8..11: T and d_t share one stacked tile
            double h0 = 0.0, h1 = 0.0, a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            dmma884(h0, h1, mh[0], d0);
            dmma884(a0, a1, mj[2], s0);
            dmma884(b0, b1, mj[2], r0);
            dmma884(h0, h1, mh[1], d1);
            dmma884(a0, a1, mj[3], s1);
            dmma884(b0, b1, mj[3], r1);
            st2((g < 4 ? arr + 8 * PS : dt + 4 * PS) + so, h0, h1);   // g >= 4: plane 8 + r = pi(g) + 4
            if (g < 4) {
              st2(ds + so + 8 * PS, a0, a1);
              st2(dr + so + 8 * PS, b0, b1);
            }
          }
        }
        {                                                             // adjoint velocity: T only
          double d0[3], d1[3];
#pragma unroll
          for (int fo = 0; fo < 3; fo++) { d0[fo] = sm[fo * AS + lo]; d1[fo] = sm[fo * AS + lo + 4 * PS]; }
          __syncwarp();
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            double c0[3] = {0.0, 0.0, 0.0}, c1[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int fo = 0; fo < 3; fo++) dmma884(c0[fo], c1[fo], mj[mt * 2], d0[fo]);
#pragma unroll
            for (int fo = 0; fo < 3; fo++) dmma884(c0[fo], c1[fo], mj[mt * 2 + 1], d1[fo]);
            if (mt == 0 || g < 4) {
#pragma unroll
              for (int fo = 0; fo < 3; fo++) st2(sm + fo * AS + so + 8 * PS * mt, c0[fo], c1[fo]);
            }
          }
        }
      }
      __syncwarp();

      // point-wise stage on the tile (in place: R -> TV, Fr -> DR, Fs -> DS, Ft -> DT).  lane -> column lane % 8, planes
      // 4 r + (0, 2, 1, 3)[lane / 8]: the two planes of a half warp are 2 apart (296 words = 8 mod 16 banks).
      // Tried and dropped: (a) handing R, Fr, Fs, Ft to T1 in registers (lane (g, q) taking the points of T1's data
      // fragments) saves 24 shared-memory accesses per point but interleaves DFMA and DMMA in every warp; the two share
      // the FP64 pipe and alternate badly (profiles/r01c: 27 instead of 58 FMA/clk/SM): 3.48 instead of 3.06 ms -- the
      // phases stay apart; (b) two adjacent columns per lane with 128-bit accesses (half the LSU instructions, but two
      // rounds with a third of the lanes idle in the second, and spills at the 96 registers of 18 warps): 3.69 ms.
      {
        const int col = 8 * t + (lane & 7), lg = lane >> 3;
        const int pl0 = ((lg & 1) << 1) | (lg >> 1);
        const int pb = col / 12, pa = col - 12 * pb;
        const double wab = W[pa] * W[pb];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const int c = 4 * r + pl0;
          double* P = sm + col + PS * c;
          double gg[9];
#pragma unroll
          for (int x = 0; x < 9; x++) gg[x] = __ldg(p.G[x] + ebd + col + PL * c);
          const double w3 = wab * W[c];
          double tv[3], tb[3];
#pragma unroll
          for (int x = 0; x < 3; x++) { tv[x] = w3 * P[x * AS]; tb[x] = P[(3 + x) * AS]; }
          double R0 = 0.0, R1 = 0.0, R2 = 0.0;
#pragma unroll
          for (int x = 0; x < 3; x++) {
            const double dr = P[(6 + x) * AS], ds = P[(9 + x) * AS], dt = P[(12 + x) * AS];
            const double dx = gg[0] * dr + gg[1] * ds + gg[2] * dt;
            const double dy = gg[3] * dr + gg[4] * ds + gg[5] * dt;
            const double dz = gg[6] * dr + gg[7] * ds + gg[8] * dt;
            R0 = fma(tv[x], dx, R0);
            R1 = fma(tv[x], dy, R1);
            R2 = fma(tv[x], dz, R2);
          }
          const double cr = tb[0] * gg[0] + tb[1] * gg[3] + tb[2] * gg[6];
          const double cs = tb[0] * gg[1] + tb[1] * gg[4] + tb[2] * gg[7];
          const double ct = tb[0] * gg[2] + tb[1] * gg[5] + tb[2] * gg[8];
          P[0 * AS] = R0; P[1 * AS] = R1; P[2 * AS] = R2;
#pragma unroll
          for (int x = 0; x < 3; x++) {
            P[(6 + x) * AS] = tv[x] * cr;
            P[(9 + x) * AS] = tv[x] * cs;
            P[(12 + x) * AS] = tv[x] * ct;
          }
        }
      }
      __syncwarp();
    }

    double pv0 = 0.0, pv1 = 0.0, pv2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;    // this element's GLL values for the epilogue
    {
      double jt[3], djt[3];    // J^T, DJ^T as A operand with output row pi(g)
#pragma unroll
      for (int ks = 0; ks < 3; ks++) { jt[ks] = FT[(8 + ks) * 32 + lane]; djt[ks] = FT[(11 + ks) * 32 + lane]; }

      // ---- T1 on the warp's column tile: t axis, K = 12; output n = pi(g): matrix as A.  X1 -> R, X2 -> Fr, X3 -> Fs
      {
        const int t = warp;
        const int lo = 8 * t + g + PS * q;
        const int so = 8 * t + 2 * q + PS * pg;
#pragma unroll 1
        for (int c = 0; c < 3; c++) {
          double* R = sm + c * AS;
          double* Fr = sm + (6 + c) * AS;
          double* Fs = sm + (9 + c) * AS;
          const double* Ft = sm + (12 + c) * AS;
          double x[3], y[3], u[3], w[3];
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            x[ks] = R[lo + 4 * PS * ks]; y[ks] = Ft[lo + 4 * PS * ks];
            u[ks] = Fr[lo + 4 * PS * ks]; w[ks] = Fs[lo + 4 * PS * ks];
          }
          __syncwarp();
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, h0 = 0.0, h1 = 0.0, k0 = 0.0, k1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            dmma884(c0, c1, jt[ks], x[ks]);
            dmma884(e0, e1, djt[ks], y[ks]);
            dmma884(h0, h1, jt[ks], u[ks]);
            dmma884(k0, k1, jt[ks], w[ks]);
          }
          st2(R + so, c0 + e0, c1 + e1);
          st2(Fr + so, h0, h1);
          st2(Fs + so, k0, k1);
        }
      }
      __syncthreads();
      // epilogue inputs of this element and F1's fragments of the next one travel during T2 / T3 (pinned: the compiler
      // would sink them)
      if (gth) {
        const size_t gi = eb + tid;
        if (flags & FLAG_ACCUM) {
#pragma unroll
          for (int c = 0; c < 3; c++) ep_f[c] = ld_pinned_rw(p.f[c] + gi);
        } else {
          if (flags & (FLAG_SOURCES | FLAG_FSTATIC)) ep_bm = ldg_pinned(p.B + gi);
          if (flags & FLAG_SOURCES) ep_rho = ldg_pinned(p.rho + gi);
          if (flags & (FLAG_SOURCES | FLAG_SENS)) {
            pv0 = ldg_pinned(p.v[0] + gi); pv1 = ldg_pinned(p.v[1] + gi); pv2 = ldg_pinned(p.v[2] + gi);
            b0 = ldg_pinned(p.vb[0] + gi); b1 = ldg_pinned(p.vb[1] + gi); b2 = ldg_pinned(p.vb[2] + gi);
          }
        }
      }
      if (it + (int)gridDim.x < p.nelem) load_f1_fragments((size_t)elem_of(it + gridDim.x) * N);

      // ---- T2: s axis, K = 12.  batch beta = i + 12 n (96); output m = pi(g).  part 0: Y1 -> R, part 1: Y2 -> Fr -----
      // 36 (component, tile) pairs x {Y1: 6 DMMAs, Y2: 3}: every warp takes two pairs whole (18 DMMAs each warp), all
      // fragment loads first
      {
        double x[2][3], y[2][3], u[2][3];
        int lo[2], so[2], cc[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int idx = warp + C::NWARP * j, c = idx / 12, t = idx - 12 * c;
          const int bl = 8 * t + g, bs = 8 * t + 2 * q;
          cc[j] = c;
          lo[j] = (bl % 12) + PS * (bl / 12) + 12 * q;               // + 48 per k-step
          so[j] = (bs % 12) + PS * (bs / 12) + 12 * pg;
          const double* R = sm + c * AS;
          const double* X2 = sm + (6 + c) * AS;
          const double* X3 = sm + (9 + c) * AS;
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            x[j][ks] = R[lo[j] + 48 * ks]; y[j][ks] = X3[lo[j] + 48 * ks]; u[j][ks] = X2[lo[j] + 48 * ks];
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            dmma884(c0, c1, jt[ks], x[j][ks]);
            dmma884(e0, e1, djt[ks], y[j][ks]);
            dmma884(h0, h1, jt[ks], u[j][ks]);
          }
          st2(sm + cc[j] * AS + so[j], c0 + e0, c1 + e1);
          st2(sm + (6 + cc[j]) * AS + so[j], h0, h1);
        }
      }
    }
    __syncthreads();

    // ---- T3: r axis, K = 12.  batch (m, n): tile t = n, entry g = row m = pi(g); output l = 2q, 2q+1: matrix as B ------
    if (warp < 12) {
      double jt[3], djt[3];
#pragma unroll
      for (int ks = 0; ks < 3; ks++) { jt[ks] = FT[(22 + ks) * 32 + lane]; djt[ks] = FT[(25 + ks) * 32 + lane]; }
      double x[2][3], y[2][3];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int tk = warp + 12 * j, c = tk >> 3, t = tk & 7;
        const double* R = sm + c * AS + 12 * pg + PS * t;
        const double* Y2 = sm + (6 + c) * AS + 12 * pg + PS * t;
#pragma unroll
        for (int ks = 0; ks < 3; ks++) { x[j][ks] = R[4 * ks + q]; y[j][ks] = Y2[4 * ks + q]; }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int tk = warp + 12 * j, c = tk >> 3, t = tk & 7;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 3; ks++) {
          dmma884(c0, c1, x[j][ks], jt[ks]);
          dmma884(e0, e1, y[j][ks], djt[ks]);
        }
        st2(sm + c * AS + 12 * pg + PS * t + 2 * q, c0 + e0, c1 + e1);
      }
    }
    __syncthreads();

    // ---- epilogue at the GLL point of this thread ------------------------------------------------------------------------
    if (gth) {
      const double o0 = sm[0 * AS + goff], o1 = sm[1 * AS + goff], o2 = sm[2 * AS + goff];
      const size_t gi = eb + tid;
      if (flags & FLAG_ACCUM) {
        p.f[0][gi] = ep_f[0] - o0; p.f[1][gi] = ep_f[1] - o1; p.f[2][gi] = ep_f[2] - o2;
      } else {
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        if (flags & (FLAG_SOURCES | FLAG_FSTATIC | FLAG_SENS)) {
          const double bm = ep_bm;
          if (flags & FLAG_SOURCES) {
            double ch = ep_rho;
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
