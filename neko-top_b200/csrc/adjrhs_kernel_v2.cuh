// Fused adjoint-RHS element kernel for sm_100a (fp64), second generation ("v2").
//
// Same arithmetic as every fused kernel (adjrhs_common.cuh: see the operator summary and the reference citations there);
// what changes is how an SM is filled.  The first kernel ran 2 CTAs of (2 consumer warps + 1 TMA warp)
// per SM: ncu showed DRAM traffic already at the algorithmic minimum but only 1.5 warps per scheduler,
// 29 % issue utilisation and consumers waiting on plane data (profiles/README.md, r01a).  Here:
//
//  * ONE persistent CTA per SM holds NE independent "element slots".  Each slot is LX*LX consumer threads
//    (thread (i,j) owns the t-pencil (i,j,0..LX-1)) with its own work arrays, its own plane ring and its
//    own named barrier; slots never synchronise with each other, so their phases interleave and hide
//    each other's latencies (8-10 consumer warps per SM instead of 4).
//  * ONE producer warp serves all slots: a single elected lane polls the slots' "empty" mbarriers
//    round-robin (non-blocking test_wait) and streams the point-wise fields of the next k-plane with
//    1-D TMA bulk copies (cp.async.bulk -> UBLKCP) that complete on the slot's "full" mbarrier.
//    Sharing the producer returns the third of the register file the per-element TMA warps used to pin.
//  * The base flow U_b (needed as whole r-, s- and t-pencils) is no longer staged in shared memory:
//    phase A reads the three pencils straight from global memory with fully sector-efficient 128/64-bit
//    loads (first touch from DRAM/L2, the other two from L1/L2); the producer issues
//    cp.async.bulk.prefetch.L2 for the slot's next element one element ahead.  This frees 12 KB per
//    slot and removes the 4-way bank conflicts of the un-swizzled TMA image.
//  * Ring slots are compile-time positions (no constant-bank look-ups in the inner loop).
//
//  * Geometry (9 cofactors + B) comes from a private per-plane interleaved image, one bulk copy per
//    plane: 5 copies per plane instead of 14 (a single issuing lane could not feed 4-5 slots otherwise).
//
// Shared memory per slot (LX = 8): 6 work arrays (24 KB, 128-byte XOR swizzle) + NS stages x NF planes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "adjrhs_common.cuh"   // operator summary, flags, PTX helpers, wsw(), load_row/store_row

namespace b200 {

// ring field positions (compile time).  NF = 14: fused path; NF = 20: + static forcing + f-in (accumulate).
// Positions 0..9 are the geometry block (9 cofactors + B): the handle keeps a private per-plane
// interleaved image of them (built once by b200_adjrhs_set_geometry, geometry is constant during a run
// exactly like coef_t), laid out [element][k][field 0..9][j][i], so the whole block of one plane is ONE
// contiguous 10*LX*LX*8-byte bulk copy instead of ten.
enum : int { R_G = 0, R_B = 9, NGEO = 10, R_V = 10, R_RHO = 13, R_FS = 14, R_FIN = 17, NF_FUSED = 14, NF_FULL = 20 };

template <int LX>
struct KParams2 {
  double D[LX * LX];          // D(i,m) at D[i + LX*m]
  double w[LX];
  const double* ub[3];        // base flow
  const double* geom;         // packed geometry image [e][k][10][LX*LX]
  const double* pf[NF_FULL - NGEO];  // other point-wise fields by ring position - NGEO (NULL: not loaded)
  double* f[3];
  double* sens;
  double* chi_out;
  const int* elem_list;       // optional list of element ids
  int nelem;
  int n_active;               // NGEO + number of non-NULL pf entries (expect_tx bytes)
  unsigned flags;
  double f_min, f_max, q, K_lube, K_sens;
  int elem_base;              // v2: added to every element index (field pointers stay 16-byte aligned for odd LX)
  const unsigned long long* xmask;  // v3 XS: bit (j + 8k) of xmask[e]: nodes (e-1; 7,j,k) and (e; 0,j,k) are summed in the kernel
};

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp (SASS: ELECT)
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

template <int LX, int NE, int NS, int NF, int NPROD = 1>
struct V2Cfg {
  static constexpr int N = LX * LX * LX;
  static constexpr int NCONS = LX * LX;
  static constexpr int NCWARP = (NCONS + 31) / 32;
  static constexpr int NCTHR = NCWARP * 32;
  static constexpr int NTHREADS = NE * NCTHR + 32 * NPROD;      // NPROD producer warps, slots dealt round-robin
  static constexpr int PLANE = LX * LX;                    // doubles
  static constexpr int PLANE_BYTES = PLANE * 8;
  // Odd LX: a k-plane of a user field (LX*LX*8 bytes, = 8 mod 16) starts on a 16-byte boundary only for every
  // other (element, plane) pair, and TMA bulk copies need 16-byte aligned source, destination and size.  The
  // producer therefore copies the aligned WINDOW of FS = PLANE_BYTES + 8 bytes that contains the plane
  // (starting 8 bytes early when (e + k) is odd) and the consumers read at the matching 8-byte shift.  The
  // window never leaves the 16-byte granule of the plane's first/last double, so it stays inside the field's
  // allocation (at most 8 bytes of a neighbouring plane, or of the granule's tail at the very end of the
  // array, are read and ignored).  The packed geometry block (10 planes = a multiple of 80 bytes) is aligned
  // for every LX.
  static constexpr bool ODD = (LX % 2 == 1);
  static constexpr int FS = ODD ? PLANE_BYTES + 8 : PLANE_BYTES;   // bytes copied per user-field plane
  static constexpr int al(int x) { return (x + 127) & ~127; }
  static constexpr int W_BYTES = al(6 * N * 8);
  static constexpr int STAGE_BYTES = al(NGEO * PLANE_BYTES + (NF - NGEO) * FS);
  static constexpr int SLOT_BYTES = W_BYTES + NS * STAGE_BYTES;
  static constexpr int BAR_OFF = NE * SLOT_BYTES;
  static constexpr int SMEM = BAR_OFF + 8 * 2 * NE * NS + 16;
};

// global-memory pencil loads of phase A
template <int LX>
__device__ __forceinline__ void ldg_row(double (&u)[LX], const double* __restrict__ g) {
  if constexpr (LX % 2 == 0) {
#pragma unroll
    for (int m = 0; m < LX; m += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(g + m));
      u[m] = t.x; u[m + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int m = 0; m < LX; m++) u[m] = __ldg(g + m);
  }
}

template <int LX, int NE, int NS, int NF, int MAXREG, int NPROD = 1>
__global__ void __launch_bounds__(V2Cfg<LX, NE, NS, NF, NPROD>::NTHREADS, 1) __maxnreg__(MAXREG)
adjrhs_v2_kernel(const __grid_constant__ KParams2<LX> p) {
  using C = V2Cfg<LX, NE, NS, NF, NPROD>;
  constexpr int N = C::N;
  constexpr int NCONS = C::NCONS;
  constexpr int NCTHR = C::NCTHR;
  constexpr int PLANE = C::PLANE;
  constexpr bool SWZ = (LX == 8);

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  // p_full[s][st] = bars[(s*NS + st)*2], p_empty[s][st] = bars[(s*NS + st)*2 + 1]

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NE * NS; i++) { mbar_init(&bars[2 * i], 1u); mbar_init(&bars[2 * i + 1], C::NCWARP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nslots = (int)gridDim.x * NE;   // element e of iteration `it` of global slot g: e = g + it*nslots

  if (tid >= NE * NCTHR) {
    // ===================== producer warp: serves every slot ==========================================
    // The whole warp runs this loop with warp-uniform control flow; one elected lane issues the copies.
    // (A `lane == 0` branch instead makes the compiler wrap every UBLKCP in a lane-serialisation loop,
    // ~20 instructions per copy, and the single producer becomes the bottleneck of the SM.)
    // (NPROD > 1: producer warp w serves the slots s with s % NPROD == w -- one elected lane issues ~5 bulk copies
    // per plane and slot, and at small lx a single warp cannot keep NE slots fed)
    const int pw = (tid - NE * NCTHR) >> 5;
    int it[NE], kk[NE], st[NE], ph[NE], nmy[NE];
    int remaining = 0;
#pragma unroll
    for (int s = 0; s < NE; s++) {
      const int g = (int)blockIdx.x * NE + s;
      nmy[s] = (p.nelem > g) ? (p.nelem - 1 - g) / nslots + 1 : 0;
      if (NPROD > 1 && (s % NPROD) != pw) nmy[s] = 0;
      it[s] = 0; kk[s] = 0; st[s] = 0; ph[s] = 0;
      remaining += (nmy[s] > 0);
    }
    const uint32_t stage_tx = (uint32_t)(NGEO * C::PLANE_BYTES + (p.n_active - NGEO) * C::FS);
    while (remaining > 0) {
      bool progressed = false;
#pragma unroll
      for (int s = 0; s < NE; s++) {
        if (it[s] >= nmy[s]) continue;
        uint64_t* full = &bars[(s * NS + st[s]) * 2];
        const bool ok = __all_sync(0xffffffffu, mbar_test_wait(full + 1, (uint32_t)(ph[s] ^ 1)));
        if (!ok) continue;
        progressed = true;
        int e = (int)blockIdx.x * NE + s + it[s] * nslots;
        if (p.elem_list) e = __ldg(p.elem_list + e);
        e += p.elem_base;
        const size_t goff = (size_t)e * N + (size_t)kk[s] * PLANE;
        // odd LX: start the user-field window one double early when the plane is not 16-byte aligned
        const size_t uoff = C::ODD ? (goff & ~(size_t)1) : goff;
        unsigned char* dst = smem + s * C::SLOT_BYTES + C::W_BYTES + st[s] * C::STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(full, stage_tx);
          tma_load_1d(dst, p.geom + goff * NGEO, NGEO * C::PLANE_BYTES, full);
#pragma unroll
          for (int a = NGEO; a < NF; a++) {
            const double* src = p.pf[a - NGEO];
            if (src) tma_load_1d(dst + NGEO * C::PLANE_BYTES + (a - NGEO) * C::FS, src + uoff, C::FS, full);
          }
        }
        if (kk[s] == 1 && it[s] + 1 < nmy[s]) {   // L2 prefetch of the slot's next base-flow element
          int en = (int)blockIdx.x * NE + s + (it[s] + 1) * nslots;
          if (p.elem_list) en = __ldg(p.elem_list + en);
          en += p.elem_base;
          const size_t eo = (size_t)en * N;
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 3; c++)
              l2_prefetch_bulk(p.ub[c] + (C::ODD ? (eo & ~(size_t)1) : eo), (N * 8) & ~15);
          }
        }
        if (++st[s] == NS) { st[s] = 0; ph[s] ^= 1; }
        if (++kk[s] == LX) {
          kk[s] = 0;
          if (++it[s] == nmy[s]) remaining--;
        }
      }
      if (!progressed) __nanosleep(32);
    }
    return;
  }

  // ========================= consumers: slot = tid / NCTHR ===========================================
  const int slot = tid / NCTHR;
  const int lt = tid - slot * NCTHR;
  const bool active = lt < NCONS;
  const int t = active ? lt : NCONS - 1;   // padding lanes mirror the last active thread, never store
  const int ti = t % LX;
  const int tj = t / LX;
  const int lane = tid & 31;
  double* W = reinterpret_cast<double*>(smem + slot * C::SLOT_BYTES);
  const unsigned char* ring = smem + slot * C::SLOT_BYTES + C::W_BYTES;
  uint64_t* sbar = &bars[slot * NS * 2];
  const int bar_id = 1 + slot;
  const double wij = p.w[ti] * p.w[tj];
  const unsigned flags = p.flags;
  const int g0 = (int)blockIdx.x * NE + slot;
  const int n_my = (p.nelem > g0) ? (p.nelem - 1 - g0) / nslots + 1 : 0;
  int st = 0, ph = 0;

  for (int it = 0; it < n_my; it++) {
    int e = g0 + it * nslots;
    if (p.elem_list) e = p.elem_list[e];
    e += p.elem_base;
    const size_t ebase = (size_t)e * N;

    // ---- phase A: r- and s-derivatives of the base flow as whole-pencil tasks, t-pencils to registers
    double ut[3][LX];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double* __restrict__ Uc = p.ub[c] + ebase;
      double ur[LX], us[LX];
      const int rb = (tj * LX + ti) * LX;             // r-pencil: row (j = ti, k = tj), contiguous
      ldg_row<LX>(ur, Uc + rb);
#pragma unroll
      for (int m = 0; m < LX; m++) us[m] = __ldg(Uc + (tj * LX + m) * LX + ti);    // s-pencil (i = ti, k = tj)
#pragma unroll
      for (int m = 0; m < LX; m++) ut[c][m] = __ldg(Uc + (m * LX + tj) * LX + ti); // t-pencil (home)
      {
        double g[LX];
#pragma unroll
        for (int i = 0; i < LX; i++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[i + LX * m], ur[m], s);
          g[i] = s;
        }
        if (active) store_row<LX, SWZ>(W + c * N, rb, g);
      }
      {
        double g[LX];
#pragma unroll
        for (int j = 0; j < LX; j++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[j + LX * m], us[m], s);
          g[j] = s;
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < LX; j++) W[(3 + c) * N + wsw<LX>((tj * LX + j) * LX + ti)] = g[j];
        }
      }
    }
    named_bar_sync(bar_id, NCTHR);

    // ---- point-wise phase, plane by plane -----------------------------------------------------------
    double acc[3][LX];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < LX; k++) acc[c][k] = 0.0;

#pragma unroll
    for (int k = 0; k < LX; k++) {
      mbar_wait(&sbar[st * 2], (uint32_t)ph);
      const double* stage = reinterpret_cast<const double*>(ring + st * C::STAGE_BYTES);
      // user fields follow the geometry block, FS bytes apart, shifted by one double for odd (e + k)
      const double* ustage = stage + NGEO * PLANE + (C::ODD ? (int)((ebase + (size_t)k * PLANE) & 1) : 0);
      constexpr int UF = C::FS / 8;
      const int pl = tj * LX + ti;
      const int q = (k * LX + tj) * LX + ti;
      const int qs = wsw<LX>(q);
      // all ring reads of this plane first, then the stage goes back to the producer
      const double v0 = ustage[(R_V - NGEO + 0) * UF + pl], v1 = ustage[(R_V - NGEO + 1) * UF + pl],
                   v2 = ustage[(R_V - NGEO + 2) * UF + pl];
      double G[9];
#pragma unroll
      for (int a = 0; a < 9; a++) G[a] = stage[(R_G + a) * PLANE + pl];
      double Bm = 0.0, chi = 0.0, fs0 = 0.0, fs1 = 0.0, fs2 = 0.0, fi0 = 0.0, fi1 = 0.0, fi2 = 0.0;
      if (flags & (FLAG_SOURCES | FLAG_FSTATIC)) Bm = stage[R_B * PLANE + pl];
      if (flags & FLAG_SOURCES) chi = ustage[(R_RHO - NGEO) * UF + pl];
      if constexpr (NF > NF_FUSED) {
        if (flags & FLAG_FSTATIC) {
          fs0 = ustage[(R_FS - NGEO + 0) * UF + pl]; fs1 = ustage[(R_FS - NGEO + 1) * UF + pl];
          fs2 = ustage[(R_FS - NGEO + 2) * UF + pl];
        }
        if (flags & FLAG_ACCUM) {
          fi0 = ustage[(R_FIN - NGEO + 0) * UF + pl]; fi1 = ustage[(R_FIN - NGEO + 1) * UF + pl];
          fi2 = ustage[(R_FIN - NGEO + 2) * UF + pl];
        }
      }
      // the stage is in registers: order these generic-proxy reads before the async-proxy refill, then release
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&sbar[st * 2 + 1]);
      if (++st == NS) { st = 0; ph ^= 1; }

      const double b0 = ut[0][k], b1 = ut[1][k], b2 = ut[2][k];
      const double w3 = wij * p.w[k];

      // source terms, then mass matrix (adjoint_pnpn.f90:669-676)
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if (flags & FLAG_SOURCES) {
        if (flags & FLAG_RAMP) {
          if (flags & FLAG_CONVEX_UP) chi = p.f_min + (p.f_max - p.f_min) * chi * (1.0 + p.q) / (chi + p.q);
          else chi = p.f_min + (p.f_max - p.f_min) * chi / (1.0 + p.q * (1.0 - chi));
        }
        if ((flags & FLAG_CHI_OUT) && active) p.chi_out[ebase + q] = chi;
        f0 = 0.0 - v0 * chi; f1 = 0.0 - v1 * chi; f2 = 0.0 - v2 * chi;
        if (flags & FLAG_FSTATIC) { f0 += fs0; f1 += fs1; f2 += fs2; }
        if (flags & FLAG_LUBE) {
          const double ck = chi * p.K_lube;
          f0 += b0 * ck; f1 += b1 * ck; f2 += b2 * ck;
        }
        f0 *= Bm; f1 *= Bm; f2 *= Bm;
      } else if (flags & FLAG_FSTATIC) {
        f0 = fs0 * Bm; f1 = fs1 * Bm; f2 = fs2 * Bm;
      }
      if (flags & FLAG_ACCUM) { f0 += fi0; f1 += fi1; f2 += fi2; }

      // (grad U_b)^T v, weak form: opgrad then vdot3 (adv_adjoint_no_dealias.f90:165-181)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      double cr = 0.0, cs = 0.0, ct = 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double gr = W[c * N + qs];
        const double gs = W[(3 + c) * N + qs];
        double gt = 0.0;
#pragma unroll
        for (int m = 0; m < LX; m++) gt = fma(p.D[k + LX * m], ut[c][m], gt);
        const double vc = (c == 0) ? v0 : (c == 1) ? v1 : v2;
        const double bc = ut[c][k];
        s0 = fma(vc, w3 * (G[0] * gr + G[1] * gs + G[2] * gt), s0);
        s1 = fma(vc, w3 * (G[3] * gr + G[4] * gs + G[5] * gt), s1);
        s2 = fma(vc, w3 * (G[6] * gr + G[7] * gs + G[8] * gt), s2);
        // contravariant base flow (groups the three cdtp calls of :297-299)
        cr = fma(bc, G[3 * c + 0], cr);
        cs = fma(bc, G[3 * c + 1], cs);
        ct = fma(bc, G[3 * c + 2], ct);
      }
      cr *= w3; cs *= w3; ct *= w3;
      acc[0][k] += f0 - s0; acc[1][k] += f1 - s1; acc[2][k] += f2 - s2;

      // fluxes: r,s parts go back to the work arrays (in place), t part is contracted right here
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double vc = (c == 0) ? v0 : (c == 1) ? v1 : v2;
        if (active) {
          W[c * N + qs] = vc * cr;
          W[(3 + c) * N + qs] = vc * cs;
        }
        const double ft = vc * ct;
#pragma unroll
        for (int k2 = 0; k2 < LX; k2++) acc[c][k2] = fma(-p.D[k + LX * k2], ft, acc[c][k2]);
      }

      if ((flags & FLAG_SENS) && active) {
        double s = b0 * v0;
        s = fma(b1, v1, s);
        s = fma(b2, v2, s);
        s = -s;
        double l = b0 * b0;       // K_sens == 0 when the lube term is off
        l = fma(b1, b1, l);
        l = fma(b2, b2, l);
        s = fma(p.K_sens, l, s);
        p.sens[ebase + q] = s;
      }
    }
    named_bar_sync(bar_id, NCTHR);

    // ---- phase B: transposed r/s contractions of the fluxes (pencil tasks, in place) ---------------
#pragma unroll
    for (int c = 0; c < 3; c++) {
      {
        double u[LX], g[LX];
        const int rb = (tj * LX + ti) * LX;
        load_row<LX, SWZ>(u, W + c * N, rb);
#pragma unroll
        for (int i = 0; i < LX; i++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[m + LX * i], u[m], s);
          g[i] = s;
        }
        if (active) store_row<LX, SWZ>(W + c * N, rb, g);
      }
      {
        double u[LX], g[LX];
#pragma unroll
        for (int m = 0; m < LX; m++) u[m] = W[(3 + c) * N + wsw<LX>((tj * LX + m) * LX + ti)];
#pragma unroll
        for (int j = 0; j < LX; j++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[m + LX * j], u[m], s);
          g[j] = s;
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < LX; j++) W[(3 + c) * N + wsw<LX>((tj * LX + j) * LX + ti)] = g[j];
        }
      }
    }
    named_bar_sync(bar_id, NCTHR);

    // ---- final: f = acc - R_r - R_s, one coalesced store per component and plane --------------------
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double* fo = p.f[c] + ebase;
#pragma unroll
        for (int k = 0; k < LX; k++) {
          const int q = (k * LX + tj) * LX + ti;
          const int qs = wsw<LX>(q);
          fo[q] = acc[c][k] - (W[c * N + qs] + W[(3 + c) * N + qs]);
        }
      }
    }
    named_bar_sync(bar_id, NCTHR);
  }
}

}  // namespace b200
