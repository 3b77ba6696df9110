// Fused adjoint-RHS element kernel for sm_100a (fp64), third generation ("v3"): lx = 8, contractions on
// the FP64 tensor-core instruction (DMMA, mma.sync.m8n8k4.f64).
//
// Why (profiles/README.md r01b + tools/fp64_pipes.cu): v2 moves only the algorithmic bytes but is bound by
// instruction issue / latency -- 500 thread-instructions per DOF (2.4x the 212 DFMA minimum), 248 registers,
// 7 warps per SM.  On B200 DMMA.8x8x4 shares the fp64 pipe with DFMA (64 FMA/clk/SM either way, measured), so
// it adds no flops -- but one DMMA issues 256 FMAs, and, more important, the mma fragment layouts remove
// almost all data movement between the contractions and the point-wise stage:
//
//   lane L of a warp: g = L>>2, q = L&3.  C fragment of m8n8k4 = C[row g][cols 2q,2q+1].
//   A warp owns a whole k-plane; lane (g,q) owns the two points (i = 2q,2q+1 ; j = g ; k).  Then
//   * r-derivative (C[j][i] = sum_m U(m,j) D(i,m)): the A fragment is the lane's own 128-bit load of U
//     (contraction index permuted m = 2q+s, s = step), B fragment is a constant piece of D, and the result
//     lands on the lane's own two points;
//   * s-derivative (C[j][i] = sum_m D(j,m) U(i,m)): A fragment constant, B fragment = two 64-bit loads of the
//     same plane (L1 hits), result again on the lane's own points;
//   * t-derivative: done per j-slab (C[k][i] = sum_m D(k,m) U(i,j,m)), the one stage that crosses planes;
//     it goes through a 12 KB swizzled shared array Wt;
//   * point-wise stage on registers (+ the plane's 14 TMA-staged fields, 128-bit conflict-free reads);
//   * transposed r contraction: the A fragment IS the lane's flux pair; transposed s needs a warp-private
//     transposition (1.5 KB scratch); both accumulate straight onto the point-wise part in the C fragment;
//   * transposed t again per j-slab through Wt; the final f = C + R_t is stored 128-bit from the fragment.
//   Per element: 288 DMMA + ~1.5 k other warp-instructions instead of ~8 k; ~100 registers; r/s work arrays gone.
//
// CTA = NE element slots x NW warps, no producer warp: the register file is split per SM sub-partition, so a
// 17th warp would cost every thread a fifth of its registers.  A plane stage is consumed by exactly one warp
// (the plane's owner), so each warp runs its own ring of DS stages: wait on the stage's mbarrier, pull the 14
// fields into registers, fence.proxy.async, and an elected lane re-arms the stage with the TMA bulk copies
// (packed geometry image + point-wise fields) of the plane it will need DS planes later.
// Arithmetic = adjrhs_common.cuh header.
//
// XS ("x stage" of the staged direct-stiffness summation, gs_kernels.cuh "staged", profiles/README.md r02): the
// class-list pass re-reads and re-writes ALL of f because every 32-byte sector of an element holds a node of an
// i-face (i = 0 or 7).  With XS each slot processes a contiguous run of elements, and for every node (0,j,k) of
// element e whose bit (j + 8k) is set in p.xmask[e] -- proven at set-up to pair with node (7,j,k) of element e-1,
// the slot's previous element -- the pair is summed here: lane (g,0) fetches the partner value (stored one
// iteration earlier by lane (g,3), an L2 hit) with an asynchronous copy into the warp's scratch, adds it to its
// own i = 0 value before the regular 128-bit store and rewrites the partner with the same sum (8-byte store into a
// line that is still in L2: the kernel's DRAM traffic does not change).  a + b is commutative, so both copies
// hold identical bits.  The y / z face passes and the class-list pass that follow never touch an i-face sector
// for its own sake any more.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "adjrhs_kernel_v2.cuh"

namespace b200 {

template <int NE, int NW, int DS, int NF>
struct V3Cfg {
  static constexpr int LX = 8;
  static constexpr int N = 512;
  static constexpr int NSLOT_THR = NW * 32;
  static constexpr int NTHREADS = NE * NSLOT_THR;
  static constexpr int PLANE_BYTES = 512;
  static constexpr int WT_BYTES = 3 * N * 8;                 // t-direction work arrays
  static constexpr int SCR_BYTES = NW * 3 * PLANE_BYTES;     // warp-private transposition scratch
  static constexpr int STAGE_BYTES = NF * PLANE_BYTES;
  static constexpr int SLOT_BYTES = WT_BYTES + SCR_BYTES + NW * DS * STAGE_BYTES;
  static constexpr int BAR_OFF = NE * SLOT_BYTES;
  // per-lane constants (fragments of D, quadrature weights) live in shared memory, [value][lane]: read from the
  // constant bank with a lane-dependent index they cost one ADU pass per distinct address (32 per warp), and the
  // compiler re-loads them every iteration instead of keeping 20 registers -- ncu r02e: the ADU pipe was the
  // busiest unit of the kernel (54 %) before this table
  static constexpr int CT_OFF = BAR_OFF + 8 * NE * NW * DS + 16;
  static constexpr int CT_VALS = 10;
  static constexpr int SMEM = CT_OFF + CT_VALS * 32 * 8;
  static_assert(CT_OFF % 16 == 0, "constant table must be 16-byte aligned");
  static_assert(8 % NW == 0, "NW must divide 8");
};

// one DMMA: C(8x8) += A(8x4, row) * B(4x8, col); lane holds A[g][q], B[q][g], C[g][2q..2q+1]
__device__ __forceinline__ void dmma(double2& c, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c.x), "+d"(c.y)
               : "d"(a), "d"(b));
}

// Wt swizzle: double index of point (i,j,k) of a t-direction work array.  16-byte chunk index
// (i>>1) + 4j + 32k with bit 2 ^= k bit 0 and bit 1 ^= k bit 1: conflict-free for 128-bit plane access
// (fixed k), 128-bit slab stores (fixed j, k = g) and 64-bit slab loads (fixed j, i = g, k = q + 4s).
__device__ __forceinline__ int wt_off(int i, int j, int k) {
  const int chunk = ((i >> 1) + 4 * j + 32 * k) ^ ((k & 1) << 2) ^ (k & 2);
  return 2 * chunk + (i & 1);
}
// warp scratch swizzle (one plane): chunk (i>>1) + 4j with bit 1 ^= j bit 1
__device__ __forceinline__ int scr_off(int i, int j) {
  const int chunk = ((i >> 1) + 4 * j) ^ (j & 2);
  return 2 * chunk + (i & 1);
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// element -> slot map of the XS kernels, shared with the set-up (capi.cu build_xstage): slot s of nslots owns
// the contiguous run [xs_run_begin(s), xs_run_begin(s + 1)) -- balanced to +-1 element.  (Windows of shorter
// runs per slot were measured too, r02c/r02f: no faster, and they link fewer faces.)
__host__ __device__ __forceinline__ int xs_run_begin(int s, int nelem, int nslots) {
  return (int)(((long long)s * nelem) / nslots);
}
// true if element e is the first of a run (its predecessor e-1 belongs to another slot)
__host__ __device__ __forceinline__ bool xs_is_run_start(int e, int nelem, int nslots) {
  const long long sidx = ((long long)e * nslots + nelem - 1) / nelem;      // smallest s with run_begin(s) >= e
  return (sidx * nelem) / nslots == e;
}

// LIST: elements come from p.elem_list (boundary / interior split of a multi-GPU run); XS: x stage (see the header)
template <int NE, int NW, int DS, int NF, int MAXREG, bool LIST = false, int XS = 0>
__global__ void __launch_bounds__(V3Cfg<NE, NW, DS, NF>::NTHREADS, 1) __maxnreg__(MAXREG)
adjrhs_v3_kernel(const __grid_constant__ KParams2<8> p) {
  using C = V3Cfg<NE, NW, DS, NF>;
  // XS != 0: lane (g,0) re-loads the previous element's i = 7 value from L2 and rewrites it -- no cross-lane
  // traffic.  (Tried and dropped, r02c/r02j: the value kept in shared memory or in registers of lane (g,3) and
  // swapped with shfl.xor(3): 4-5 % slower.)
  static_assert(!XS || !LIST, "the x stage runs on contiguous element runs");
  constexpr int LX = 8, N = 512, PLANE = 64;
  constexpr int NPL = LX / NW;        // planes per warp
  constexpr int NTASK = 24;           // (component, j-slab) tasks of the t-direction stages

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NE * NW * DS; i++) mbar_init(&bars[i], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    double* ct = reinterpret_cast<double*>(smem + C::CT_OFF) + tid;
    const int tg = tid >> 2, tq = tid & 3;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      ct[(0 + s) * 32] = p.D[tg + 8 * (2 * tq + s)];       // B frag, forward r:   D(i=g, m=2q+s)
      ct[(2 + s) * 32] = p.D[tg + 8 * (tq + 4 * s)];       // A frag, forward s/t: D(row=g, m=q+4s)
      ct[(4 + s) * 32] = p.D[(2 * tq + s) + 8 * tg];       // B frag, transposed r:   D(m=2q+s, i=g)
      ct[(6 + s) * 32] = p.D[(tq + 4 * s) + 8 * tg];       // A frag, transposed s/t: D(m=q+4s, row=g)
      ct[(8 + s) * 32] = p.w[2 * tq + s] * p.w[tg];        // w_i w_j of the lane's points
    }
  }
  __syncthreads();

  const int nslots = (int)gridDim.x * NE;

  // ========================= consumers ================================================================
  const int slot = tid / C::NSLOT_THR;
  const int wid = (tid - slot * C::NSLOT_THR) >> 5;   // warp inside the slot
  const int lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  unsigned char* sbase = smem + slot * C::SLOT_BYTES;
  double* Wt = reinterpret_cast<double*>(sbase);
  double* scr = reinterpret_cast<double*>(sbase + C::WT_BYTES + wid * 3 * C::PLANE_BYTES);
  unsigned char* ring = sbase + C::WT_BYTES + C::SCR_BYTES + wid * DS * C::STAGE_BYTES;   // this warp's stages
  uint64_t* wbar = &bars[(slot * NW + wid) * DS];
  const int bar_id = 1 + slot;
  const unsigned flags = p.flags;

  // constant fragments of the derivative matrix D(i,m) = p.D[i + 8m] and the (i,j) quadrature weights of the
  // lane's two points: table in shared memory, filled once by the first warp
  enum : int { CT_DRF = 0, CT_DSF = 2, CT_DRB = 4, CT_DSB = 6, CT_WIJ = 8 };
  const double* ctab = reinterpret_cast<const double*>(smem + C::CT_OFF) + lane;
#define CT(v) ctab[(v) * 32]

  // per-lane static offsets
  const int pl2 = 2 * q + 8 * g;                 // own point pair inside a plane (doubles)
  const int sA0 = g + 8 * q, sA1 = g + 8 * (q + 4);   // s-direction B-fragment loads inside a plane
  const int scr_w = scr_off(2 * q, g);
  const int scr_r0 = scr_off(g, q), scr_r1 = scr_off(g, q + 4);

  const int g0 = (int)blockIdx.x * NE + slot;
  // XS: slot g0 owns the contiguous run [e_first, e_first + n_my); otherwise the elements g0, g0 + nslots, ...
  // (the grid works on a window of nslots consecutive elements)
  [[maybe_unused]] int e_first = 0;
  int n_my_;
  if constexpr (XS) {
    e_first = xs_run_begin(g0, p.nelem, nslots);
    n_my_ = xs_run_begin(g0 + 1, p.nelem, nslots) - e_first;
  } else {
    n_my_ = (p.nelem > g0) ? (p.nelem - 1 - g0) / nslots + 1 : 0;
  }
  const int n_my = n_my_;
  const int n_planes = n_my * NPL;               // planes this warp will consume
  const uint32_t stage_tx = (uint32_t)p.n_active * C::PLANE_BYTES;

  // (re-)arm stage n % DS with the n-th plane of this warp: element iteration n / NPL, plane wid + (n % NPL) NW
  // element ids of the current and the next iteration of this slot (the list look-up stays off the
  // critical path of re-arming a stage)
  // (LIST = false: elements g0 + it*nslots, no look-up and no extra registers)
  auto elem_at = [&](int itx) {
    int en = g0 + itx * nslots;
    if constexpr (LIST) { if (itx < n_my) en = __ldg(p.elem_list + en); }
    return en;
  };
  [[maybe_unused]] int it_now = 0;
  [[maybe_unused]] int e_cur = 0, e_nxt = 0;
  if constexpr (LIST) { e_cur = elem_at(0); e_nxt = elem_at(1); }
  auto issue = [&](int n) {
    if (n >= n_planes) return;
    const int itn = n / NPL, pin = n - itn * NPL;      // itn is it_now or it_now + 1 (DS <= NPL)
    int en;
    if constexpr (LIST) en = (itn == it_now) ? e_cur : e_nxt;
    else if constexpr (XS) en = e_first + itn;
    else en = g0 + itn * nslots;
    const size_t goff = (size_t)en * N + (size_t)(wid + pin * NW) * PLANE;
    unsigned char* dst = ring + (n % DS) * C::STAGE_BYTES;
    uint64_t* full = &wbar[n % DS];
    if (elect_one()) {
      mbar_expect_tx(full, stage_tx);
      tma_load_1d(dst, p.geom + goff * NGEO, NGEO * C::PLANE_BYTES, full);
#pragma unroll
      for (int a = NGEO; a < NF; a++) {
        const double* src = p.pf[a - NGEO];
        if (src) tma_load_1d(dst + a * C::PLANE_BYTES, src + goff, C::PLANE_BYTES, full);
      }
    }
  };
#pragma unroll
  for (int n = 0; n < DS; n++) issue(n);

  for (int it = 0; it < n_my; it++) {
    int e;
    if constexpr (LIST) { it_now = it; e = e_cur; }
    else if constexpr (XS) e = e_first + it;
    else e = g0 + it * nslots;
    const size_t ebase = (size_t)e * N;
    // xmask[e] bit (j + 8k): node (0,j,k) of e and node (7,j,k) of element e-1 are summed here; e-1 is then the
    // element this slot processed in its previous iteration (the set-up knows the runs: xs_is_run_start)
    [[maybe_unused]] unsigned long long xm = 0ull;
    if constexpr (XS) xm = __ldg(p.xmask + e);

    // the s/t fragments of D are read from the table once per element (they feed 16 DMMA pairs each); the r
    // fragments and the weights, used once per plane, are read where they are needed
    const double dsf0 = CT(CT_DSF), dsf1 = CT(CT_DSF + 1), dsb0 = CT(CT_DSB), dsb1 = CT(CT_DSB + 1);

    // ---- t-derivatives of the base flow, per (component, j) slab -> Wt --------------------------------
#pragma unroll
    for (int tk = 0; tk < NTASK / NW; tk++) {
      const int task = wid + tk * NW;
      const int c = task >> 3, j = task & 7;
      const double* __restrict__ Uc = p.ub[c] + ebase + 8 * j + g;
      const double b0 = __ldg(Uc + 64 * q), b1 = __ldg(Uc + 64 * (q + 4));
      double2 acc = make_double2(0.0, 0.0);
      dmma(acc, dsf0, b0);
      dmma(acc, dsf1, b1);
      *reinterpret_cast<double2*>(Wt + c * N + wt_off(2 * q, j, g)) = acc;
    }
    named_bar_sync(bar_id, C::NSLOT_THR);

    // ---- per owned plane: r/s derivatives, point-wise stage, transposed r/s ---------------------------
    double2 cacc[NPL][3];
#pragma unroll
    for (int pi = 0; pi < NPL; pi++) {
      const int k = wid + pi * NW;
      const int qoff = 64 * k + pl2;
      double2 ub2[3], gr[3], gs[3], gt[3];
      const double drf0 = CT(CT_DRF), drf1 = CT(CT_DRF + 1);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double* __restrict__ Uc = p.ub[c] + ebase + 64 * k;
        ub2[c] = __ldg(reinterpret_cast<const double2*>(Uc + pl2));
        const double s0 = __ldg(Uc + sA0), s1 = __ldg(Uc + sA1);
        gr[c] = make_double2(0.0, 0.0);
        dmma(gr[c], ub2[c].x, drf0);
        dmma(gr[c], ub2[c].y, drf1);
        gs[c] = make_double2(0.0, 0.0);
        dmma(gs[c], dsf0, s0);
        dmma(gs[c], dsf1, s1);
        gt[c] = *reinterpret_cast<const double2*>(Wt + c * N + wt_off(2 * q, g, k));
      }

      // plane stage of the ring
      const int np = it * NPL + pi;               // running plane number of this warp
      const int st = np % DS;
      const uint32_t ph = (uint32_t)(np / DS) & 1u;
      mbar_wait(&wbar[st], ph);
      const double2* stage = reinterpret_cast<const double2*>(ring + st * C::STAGE_BYTES) + lane;
      // (field a of the stage at stage[a*32])
      const double2 v0 = stage[(R_V + 0) * 32], v1 = stage[(R_V + 1) * 32], v2 = stage[(R_V + 2) * 32];
      double2 G[9];
#pragma unroll
      for (int a = 0; a < 9; a++) G[a] = stage[(R_G + a) * 32];
      double2 Bm = make_double2(0.0, 0.0), chi = Bm, fs0 = Bm, fs1 = Bm, fs2 = Bm, fi0 = Bm, fi1 = Bm, fi2 = Bm;
      if (flags & (FLAG_SOURCES | FLAG_FSTATIC)) Bm = stage[R_B * 32];
      if (flags & FLAG_SOURCES) chi = stage[R_RHO * 32];
      if constexpr (NF > NF_FUSED) {
        if (flags & FLAG_FSTATIC) { fs0 = stage[(R_FS + 0) * 32]; fs1 = stage[(R_FS + 1) * 32]; fs2 = stage[(R_FS + 2) * 32]; }
        if (flags & FLAG_ACCUM) { fi0 = stage[(R_FIN + 0) * 32]; fi1 = stage[(R_FIN + 1) * 32]; fi2 = stage[(R_FIN + 2) * 32]; }
      }
      // the stage is in registers: order the generic-proxy reads before the async-proxy refill, re-arm it
      fence_proxy_async_smem();
      __syncwarp();
      issue(np + DS);
      if (pi == 0 && wid == 0 && it + 1 < n_my) {   // L2 prefetch of the slot's next base-flow element
        int en;
        if constexpr (LIST) en = e_nxt;
        else if constexpr (XS) en = e + 1;
        else en = g0 + (it + 1) * nslots;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 3; c++) l2_prefetch_bulk(p.ub[c] + (size_t)en * N, N * 8);
        }
      }

      const double wk = p.w[k];
      double2 fpw[3], Fr[3], Fs[3], Ft[3], sens2;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        auto sel = [h](const double2& x) { return h ? x.y : x.x; };
        const double w3 = CT(CT_WIJ + h) * wk;
        const double pv0 = sel(v0), pv1 = sel(v1), pv2 = sel(v2);
        const double b0 = sel(ub2[0]), b1 = sel(ub2[1]), b2 = sel(ub2[2]);
        double Gp[9];
#pragma unroll
        for (int a = 0; a < 9; a++) Gp[a] = sel(G[a]);
        // source terms, then mass matrix (adjoint_pnpn.f90:669-676)
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        double ch = sel(chi);
        const double bm = sel(Bm);
        if (flags & FLAG_SOURCES) {
          if (flags & FLAG_RAMP) {
            if (flags & FLAG_CONVEX_UP) ch = p.f_min + (p.f_max - p.f_min) * ch * (1.0 + p.q) / (ch + p.q);
            else ch = p.f_min + (p.f_max - p.f_min) * ch / (1.0 + p.q * (1.0 - ch));
          }
          f0 = 0.0 - pv0 * ch; f1 = 0.0 - pv1 * ch; f2 = 0.0 - pv2 * ch;
          if (flags & FLAG_FSTATIC) { f0 += sel(fs0); f1 += sel(fs1); f2 += sel(fs2); }
          if (flags & FLAG_LUBE) {
            const double ck = ch * p.K_lube;
            f0 += b0 * ck; f1 += b1 * ck; f2 += b2 * ck;
          }
          f0 *= bm; f1 *= bm; f2 *= bm;
        } else if (flags & FLAG_FSTATIC) {
          f0 = sel(fs0) * bm; f1 = sel(fs1) * bm; f2 = sel(fs2) * bm;
        }
        if (flags & FLAG_ACCUM) { f0 += sel(fi0); f1 += sel(fi1); f2 += sel(fi2); }
        if (h) chi.y = ch; else chi.x = ch;

        // (grad U_b)^T v, weak form (adv_adjoint_no_dealias.f90:165-181) and contravariant base flow
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, cr = 0.0, cs = 0.0, ct = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double dr = sel(gr[c]), ds = sel(gs[c]), dt = sel(gt[c]);
          const double vc = (c == 0) ? pv0 : (c == 1) ? pv1 : pv2;
          const double bc = (c == 0) ? b0 : (c == 1) ? b1 : b2;
          s0 = fma(vc, w3 * (Gp[0] * dr + Gp[1] * ds + Gp[2] * dt), s0);
          s1 = fma(vc, w3 * (Gp[3] * dr + Gp[4] * ds + Gp[5] * dt), s1);
          s2 = fma(vc, w3 * (Gp[6] * dr + Gp[7] * ds + Gp[8] * dt), s2);
          cr = fma(bc, Gp[3 * c + 0], cr);
          cs = fma(bc, Gp[3 * c + 1], cs);
          ct = fma(bc, Gp[3 * c + 2], ct);
        }
        cr *= -w3; cs *= -w3; ct *= -w3;     // negated fluxes: f = fpw + D^T(-flux)
        const double p0 = f0 - s0, p1 = f1 - s1, p2 = f2 - s2;
        double sv = b0 * pv0;
        sv = fma(b1, pv1, sv);
        sv = fma(b2, pv2, sv);
        sv = -sv;
        double l = b0 * b0;       // K_sens == 0 when the lube term is off
        l = fma(b1, b1, l);
        l = fma(b2, b2, l);
        sv = fma(p.K_sens, l, sv);
        if (h) {
          fpw[0].y = p0; fpw[1].y = p1; fpw[2].y = p2; sens2.y = sv;
          Fr[0].y = pv0 * cr; Fr[1].y = pv1 * cr; Fr[2].y = pv2 * cr;
          Fs[0].y = pv0 * cs; Fs[1].y = pv1 * cs; Fs[2].y = pv2 * cs;
          Ft[0].y = pv0 * ct; Ft[1].y = pv1 * ct; Ft[2].y = pv2 * ct;
        } else {
          fpw[0].x = p0; fpw[1].x = p1; fpw[2].x = p2; sens2.x = sv;
          Fr[0].x = pv0 * cr; Fr[1].x = pv1 * cr; Fr[2].x = pv2 * cr;
          Fs[0].x = pv0 * cs; Fs[1].x = pv1 * cs; Fs[2].x = pv2 * cs;
          Ft[0].x = pv0 * ct; Ft[1].x = pv1 * ct; Ft[2].x = pv2 * ct;
        }
      }
      if (flags & FLAG_SENS) *reinterpret_cast<double2*>(p.sens + ebase + qoff) = sens2;
      if (flags & FLAG_CHI_OUT) *reinterpret_cast<double2*>(p.chi_out + ebase + qoff) = chi;

      // t fluxes back to Wt (same locations this lane read gt from), s fluxes to the warp scratch
#pragma unroll
      for (int c = 0; c < 3; c++) {
        *reinterpret_cast<double2*>(Wt + c * N + wt_off(2 * q, g, k)) = Ft[c];
        *reinterpret_cast<double2*>(scr + c * PLANE + scr_w) = Fs[c];
      }
      __syncwarp();
      const double drb0 = CT(CT_DRB), drb1 = CT(CT_DRB + 1);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double2 acc = fpw[c];
        dmma(acc, Fr[c].x, drb0);                          // transposed r: A = own flux pair
        dmma(acc, Fr[c].y, drb1);
        const double t0 = scr[c * PLANE + scr_r0], t1 = scr[c * PLANE + scr_r1];
        dmma(acc, dsb0, t0);                               // transposed s
        dmma(acc, dsb1, t1);
        cacc[pi][c] = acc;
      }
      __syncwarp();   // scratch is reused by the next plane
    }
    named_bar_sync(bar_id, C::NSLOT_THR);

    // ---- x stage, part 1: fetch the previous element's i = 7 values (L2 hits, ~1 us under load) as asynchronous
    //      8-byte copies (LDGSTS) into the warp's scratch, which is free between the plane loop and the next
    //      element's: no register is held across the transposed t contraction, whose work hides the latency.
    //      (The same fetch into registers: 1-1.5 % slower step, r02v; right before the stores: +7 %, r02l.)
    [[maybe_unused]] bool xdo[NPL];
    if constexpr (XS != 0) {
#pragma unroll
      for (int pi = 0; pi < NPL; pi++) {
        const int k = wid + pi * NW;
        xdo[pi] = q == 0 && ((xm >> (g + 8 * k)) & 1ull);
        if (xdo[pi]) {
          const size_t xoff = ebase - N + 64 * k + 8 * g + 7;        // (i = 7, j = g, k) of the slot's previous element
#pragma unroll
          for (int c = 0; c < 3; c++) cp_async_8(scr + (pi * 3 + c) * 8 + g, p.f[c] + xoff);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- transposed t contraction per (component, j) slab, in place in Wt ------------------------------
#pragma unroll
    for (int tk = 0; tk < NTASK / NW; tk++) {
      const int task = wid + tk * NW;
      const int c = task >> 3, j = task & 7;
      const double b0 = Wt[c * N + wt_off(g, j, q)], b1 = Wt[c * N + wt_off(g, j, q + 4)];
      double2 acc = make_double2(0.0, 0.0);
      dmma(acc, dsb0, b0);
      dmma(acc, dsb1, b1);
      __syncwarp();
      *reinterpret_cast<double2*>(Wt + c * N + wt_off(2 * q, j, g)) = acc;
    }
    named_bar_sync(bar_id, C::NSLOT_THR);

    // ---- final: f = (fpw + R_r + R_s) + R_t, 128-bit stores straight from the fragments -----------------
    [[maybe_unused]] double pv[NPL][3];
    if constexpr (XS != 0) {     // x stage, part 2: the lane that issued a copy reads it back (its own copies only)
      asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
      for (int pi = 0; pi < NPL; pi++)
#pragma unroll
        for (int c = 0; c < 3; c++) pv[pi][c] = xdo[pi] ? scr[(pi * 3 + c) * 8 + g] : 0.0;
    }
#pragma unroll
    for (int pi = 0; pi < NPL; pi++) {
      const int k = wid + pi * NW;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double2 rt = *reinterpret_cast<const double2*>(Wt + c * N + wt_off(2 * q, g, k));
        double2 o;
        o.x = cacc[pi][c].x + rt.x;
        o.y = cacc[pi][c].y + rt.y;
        if constexpr (XS != 0) {
          if (xdo[pi]) { o.x += pv[pi][c]; p.f[c][ebase - N + 64 * k + 8 * g + 7] = o.x; }
        }
        *reinterpret_cast<double2*>(p.f[c] + ebase + 64 * k + pl2) = o;
      }
    }
    named_bar_sync(bar_id, C::NSLOT_THR);

    if constexpr (LIST) {
      e_cur = e_nxt;
      e_nxt = elem_at(it + 2);
    }
  }
#undef CT
}

}  // namespace b200
