// Fused adjoint-RHS element kernel for sm_100a (fp64).
//
// One pass over each hexahedral element computes what the reference does in ~45 whole-field
// sweeps (SURVEY.md 8a "fused operator"; citations relative to /root/reference/sources):
//   chi  = RAMP(rho)                                      mapping_functions/RAMP_mapping.f90:227-241
//   f_i  = B*(-chi*v_i [+K*chi*vb_i] [+fs_i])             source_terms/simple_brinkman_source_term.f90:149-151,
//                                                         source_terms/adjoint_lube_source_term.f90:189-203,
//                                                         adjoint/adjoint_pnpn.f90:672-676
//   f_i -= sum_j v_j * opgrad(vb_j)_i                     adjoint/adv_adjoint_no_dealias.f90:165-181
//   f_i -= sum_k cdtp(v_i*vb_k ; d./dx_k)                 adjoint/adv_adjoint_no_dealias.f90:183-201,269-303
//   S    = -(vb.v) + K_s*(vb.vb)                          objectives/minimum_dissipation_objective_function.f90:260-301
//
// Design:
//  * persistent CTAs, grid = #SM x CTAs/SM, elements strided over CTAs (deterministic);
//  * one producer warp streams the element's fields global->shared with 1-D TMA bulk copies
//    (cp.async.bulk ... mbarrier::complete_tx) through two mbarrier rings: a "U ring" holding the
//    three base-flow components of a whole element (needed by the r/s pencil contractions) and a
//    "plane ring" holding PC k-planes of the 14..20 point-wise fields;
//  * LX*LX consumer threads per element.  Thread (i,j) is the "home" of the t-pencil (i,j,0..LX-1):
//    t-direction contractions never leave registers; r/s contractions are done as whole-pencil
//    tasks (8 loads -> 64 DFMA with the derivative matrix coming from the constant bank through
//    the __grid_constant__ parameter block -> 8 stores) on a swizzled shared work array;
//  * the three cdtp calls per component are grouped into one contravariant flux (9 instead of 27
//    transposed contractions);
//  * f and S are written once, straight from registers (256 B contiguous per warp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ---- point-wise field slots in the plane ring -------------------------------------------------
enum : int {
  PF_VX = 0, PF_VY = 1, PF_VZ = 2,      // adjoint velocity
  PF_G0 = 3,                            // drdx,dsdx,dtdx, drdy,dsdy,dtdy, drdz,dsdz,dtdz
  PF_B = 12,
  PF_RHO = 13,                          // rho (RAMP in kernel) or chi
  PF_FS0 = 14,                          // static forcing x,y,z
  PF_FIN0 = 17,                         // f in (accumulate mode) x,y,z
  PF_COUNT = 20
};

enum : unsigned {
  FLAG_SOURCES = 1u,     // Brinkman term active (rho or chi present)
  FLAG_RAMP = 2u,        // PF_RHO holds rho -> apply RAMP; else it already is chi
  FLAG_LUBE = 4u,        // + K_lube*chi*vb_i
  FLAG_FSTATIC = 8u,
  FLAG_ACCUM = 16u,      // f is in/out (un-fused advection_adjoint_t drop-in)
  FLAG_SENS = 32u,
  FLAG_CHI_OUT = 64u,
  FLAG_CONVEX_UP = 128u,
  FLAG_LINEAR = 256u,    // (reserved)
  FLAG_GS = 512u,        // v3 kernel: direct-stiffness summation inside the element kernel
  FLAG_L2HINT = 1024u    // v3 kernel: L2 evict_first policy on the streaming inputs / sens / chi
};

template <int LX>
struct KParams {
  double D[LX * LX];          // D(i,m) at D[i + LX*m]  (Xh%dx, column-major)
  double w[LX];               // Xh%wx
  const double* ub[3];        // base flow (whole element staged in the U ring)
  const double* pf[PF_COUNT]; // point-wise fields (NULL if inactive)
  int pf_slot[PF_COUNT];      // position of the field inside a plane-ring slot (-1 inactive)
  int n_pf;                   // number of active point-wise fields
  double* f[3];
  double* sens;
  double* chi_out;
  const int* elem_list;       // optional list of element ids (NULL: 0..nelem-1)
  int nelem;
  unsigned flags;
  double f_min, f_max, q, K_lube, K_sens;
};

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// same with an L2 eviction-priority hint (createpolicy handle)
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_f64x2_hint(double* addr, double a, double b, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(addr), "d"(a), "d"(b), "l"(policy)
               : "memory");
}
// 8-byte Ampere-style async copy (SASS: LDGSTS) + deferred arrive, for odd LX
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- shared-memory geometry -----------------------------------------------------------------------
template <int LX, int PC, int NS, int NU>
struct SmemLayout {
  static constexpr int N = LX * LX * LX;
  static constexpr int ARR_BYTES = N * 8;
  static constexpr int PLANE_BYTES = LX * LX * 8;
  static constexpr int CHUNK_BYTES = PC * PLANE_BYTES;     // one field, PC planes
  static constexpr int NCHUNK = LX / PC;
  static_assert(LX % PC == 0, "PC must divide LX");
  // 1-D TMA bulk copies need 16-byte aligned sources/sizes: true for even LX.  Odd LX falls back to
  // 8-byte cp.async (LDGSTS) issued by the 32 producer lanes, same mbarrier pipeline.
  static constexpr bool BULK = (CHUNK_BYTES % 16 == 0) && (ARR_BYTES % 16 == 0);
  static constexpr int al(int x) { return (x + 127) & ~127; }
  static constexpr int U_OFF = 0;                               // NU slots x 3 arrays
  static constexpr int W_OFF = al(U_OFF + NU * 3 * ARR_BYTES);  // 6 work arrays
  static constexpr int P_OFF = al(W_OFF + 6 * ARR_BYTES);       // NS slots x n_pf chunks (runtime n_pf)
  static constexpr int bar_off(int n_pf) { return al(P_OFF + NS * n_pf * CHUNK_BYTES); }
  static constexpr int total(int n_pf) { return bar_off(n_pf) + 8 * (2 * NU + 2 * NS) + 16; }
};

// 128-byte XOR swizzle of a double index inside a work array; conflict-free for the r-pencil
// (LDS.128 rows), s-pencil and t-home access patterns when LX == 8.
template <int LX>
__device__ __forceinline__ int wsw(int q) {
  if constexpr (LX == 8) {
    return q ^ (((q >> 4) & 7) << 1);
  } else {
    return q;
  }
}

// row (r-pencil) load/store helpers: 128-bit accesses when the row is 16-byte aligned (even LX)
template <int LX, bool SWZ>
__device__ __forceinline__ void load_row(double (&u)[LX], const double* arr, int rb) {
  if constexpr (LX % 2 == 0) {
    const int y = SWZ ? (((rb >> 4) & 7) << 1) : 0;   // constant inside a row (LX == 8)
#pragma unroll
    for (int m = 0; m < LX; m += 2) {
      const double2 t = *reinterpret_cast<const double2*>(arr + ((rb + m) ^ y));
      u[m] = t.x; u[m + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int m = 0; m < LX; m++) u[m] = arr[rb + m];
  }
}
template <int LX, bool SWZ>
__device__ __forceinline__ void store_row(double* arr, int rb, const double (&g)[LX]) {
  if constexpr (LX % 2 == 0) {
    const int y = SWZ ? (((rb >> 4) & 7) << 1) : 0;
#pragma unroll
    for (int m = 0; m < LX; m += 2)
      *reinterpret_cast<double2*>(arr + ((rb + m) ^ y)) = make_double2(g[m], g[m + 1]);
  } else {
#pragma unroll
    for (int m = 0; m < LX; m++) arr[rb + m] = g[m];
  }
}

template <int LX, int PC, int NS, int NU>
struct KernelCfg {
  static constexpr int NCONS = LX * LX;                // active consumer threads (one t-pencil each)
  static constexpr int NCWARP = (NCONS + 31) / 32;
  static constexpr int NCTHR = NCWARP * 32;            // consumer threads incl. padding lanes
  static constexpr int NTHREADS = NCTHR + 32;          // + producer warp
};

template <int LX, int PC, int NS, int NU, int MAXREG>
__global__ void __launch_bounds__(KernelCfg<LX, PC, NS, NU>::NTHREADS) __maxnreg__(MAXREG)
adjrhs_fused_kernel(const __grid_constant__ KParams<LX> p) {
  using L = SmemLayout<LX, PC, NS, NU>;
  using C = KernelCfg<LX, PC, NS, NU>;
  constexpr int N = L::N;
  constexpr int NCONS = C::NCONS;
  constexpr int NCTHR = C::NCTHR;
  constexpr bool SWZ = (LX == 8);

  extern __shared__ __align__(128) unsigned char smem[];
  double* Us = reinterpret_cast<double*>(smem + L::U_OFF);
  double* W = reinterpret_cast<double*>(smem + L::W_OFF);
  unsigned char* Ps = smem + L::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off(p.n_pf));
  uint64_t* u_full = bars;
  uint64_t* u_empty = bars + NU;
  uint64_t* p_full = bars + 2 * NU;
  uint64_t* p_empty = bars + 2 * NU + NS;

  const int tid = threadIdx.x;
  const int n_pf = p.n_pf;
  const int slot_bytes = n_pf * L::CHUNK_BYTES;

  if (tid == 0) {
    const uint32_t full_cnt = L::BULK ? 1u : 32u;
    for (int s = 0; s < NU; s++) { mbar_init(&u_full[s], full_cnt); mbar_init(&u_empty[s], 1); }
    for (int s = 0; s < NS; s++) { mbar_init(&p_full[s], full_cnt); mbar_init(&p_empty[s], C::NCWARP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_my = (p.nelem > (int)blockIdx.x) ? (p.nelem - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid >= NCTHR) {
    // ===================== producer warp ===========================================================
    const int lane = tid - NCTHR;
    if (!L::BULK || lane == 0) {
      int us = 0, uph = 0, ps = 0, pph = 0;
      auto elem_of = [&](int it) {
        int e = (int)blockIdx.x + it * (int)gridDim.x;
        return p.elem_list ? p.elem_list[e] : e;
      };
      auto issue_u = [&](int it) {
        const int e = elem_of(it);
        mbar_wait(&u_empty[us], uph ^ 1);
        if constexpr (L::BULK) {
          mbar_expect_tx(&u_full[us], 3 * L::ARR_BYTES);
#pragma unroll
          for (int c = 0; c < 3; c++)
            tma_load_1d(Us + (us * 3 + c) * N, p.ub[c] + (size_t)e * N, L::ARR_BYTES, &u_full[us]);
        } else {
#pragma unroll
          for (int c = 0; c < 3; c++)
            for (int x = lane; x < N; x += 32) cp_async_8(Us + (us * 3 + c) * N + x, p.ub[c] + (size_t)e * N + x);
          cp_async_mbar_arrive(&u_full[us]);
        }
        if (++us == NU) { us = 0; uph ^= 1; }
      };
      if (n_my > 0) issue_u(0);
      for (int it = 0; it < n_my; it++) {
        if (NU > 1 && it + 1 < n_my) issue_u(it + 1);
        const int e = elem_of(it);
        for (int ch = 0; ch < L::NCHUNK; ch++) {
          mbar_wait(&p_empty[ps], pph ^ 1);
          unsigned char* dst = Ps + (size_t)ps * slot_bytes;
          const size_t goff = (size_t)e * N + (size_t)ch * PC * LX * LX;
          if constexpr (L::BULK) {
            mbar_expect_tx(&p_full[ps], (uint32_t)slot_bytes);
#pragma unroll 1
            for (int a = 0; a < PF_COUNT; a++) {
              const int sl = p.pf_slot[a];
              if (sl >= 0) tma_load_1d(dst + sl * L::CHUNK_BYTES, p.pf[a] + goff, L::CHUNK_BYTES, &p_full[ps]);
            }
          } else {
#pragma unroll 1
            for (int a = 0; a < PF_COUNT; a++) {
              const int sl = p.pf_slot[a];
              if (sl < 0) continue;
              double* d = reinterpret_cast<double*>(dst + sl * L::CHUNK_BYTES);
              for (int x = lane; x < PC * LX * LX; x += 32) cp_async_8(d + x, p.pf[a] + goff + x);
            }
            cp_async_mbar_arrive(&p_full[ps]);
          }
          if (++ps == NS) { ps = 0; pph ^= 1; }
        }
        if (NU == 1 && it + 1 < n_my) issue_u(it + 1);
      }
      if constexpr (!L::BULK) asm volatile("cp.async.wait_all;" ::: "memory");
    }
    return;
  }

  // ========================= consumers ===========================================================
  // padding lanes (tid >= NCONS, only when LX*LX is not a multiple of 32) mirror the last active
  // thread and never store.
  const bool active = tid < NCONS;
  const int t = active ? tid : NCONS - 1;
  const int ti = t % LX;       // home (i,j); also s-pencil (i, k=tj) and r-pencil (j=ti, k=tj)
  const int tj = t / LX;
  const int lane = tid & 31;
  const double wij = p.w[ti] * p.w[tj];
  const unsigned flags = p.flags;
  int us = 0, uph = 0, ps = 0, pph = 0;

  for (int it = 0; it < n_my; it++) {
    int e = (int)blockIdx.x + it * (int)gridDim.x;
    if (p.elem_list) e = p.elem_list[e];
    const size_t ebase = (size_t)e * N;

    // ---- phase A: r- and s-derivatives of the base flow (pencil tasks), t-pencils to registers ---
    mbar_wait(&u_full[us], uph);
    const double* U0 = Us + us * 3 * N;
    double ut[3][LX];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double* Uc = U0 + c * N;
      {  // r-pencil: row (j=ti, k=tj), contiguous
        double u[LX], g[LX];
        const int rb = (tj * LX + ti) * LX;
        load_row<LX, false>(u, Uc, rb);
#pragma unroll
        for (int i = 0; i < LX; i++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[i + LX * m], u[m], s);
          g[i] = s;
        }
        if (active) store_row<LX, SWZ>(W + c * N, rb, g);
      }
      {  // s-pencil: (i=ti, k=tj)
        double u[LX], g[LX];
#pragma unroll
        for (int m = 0; m < LX; m++) u[m] = Uc[(tj * LX + m) * LX + ti];
#pragma unroll
        for (int j = 0; j < LX; j++) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < LX; m++) s = fma(p.D[j + LX * m], u[m], s);
          g[j] = s;
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < LX; j++) W[(3 + c) * N + wsw<LX>((tj * LX + j) * LX + ti)] = g[j];
        }
      }
      // t-pencil (home)
#pragma unroll
      for (int m = 0; m < LX; m++) ut[c][m] = Uc[(m * LX + tj) * LX + ti];
    }
    named_bar_sync(1, NCTHR);
    if (tid == 0) mbar_arrive(&u_empty[us]);
    if (++us == NU) { us = 0; uph ^= 1; }

    // ---- point-wise phase, plane by plane ----------------------------------------------------------
    double acc[3][LX];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < LX; k++) acc[c][k] = 0.0;

#pragma unroll
    for (int k = 0; k < LX; k++) {
      if (k % PC == 0) mbar_wait(&p_full[ps], pph);
      const unsigned char* slot = Ps + (size_t)ps * slot_bytes;
      const int pl = (k % PC) * LX * LX + tj * LX + ti;    // index inside a chunk
      const int q = (k * LX + tj) * LX + ti;               // index inside the element
      const int qs = wsw<LX>(q);
      auto PF = [&](int a) -> double {
        return reinterpret_cast<const double*>(slot + p.pf_slot[a] * L::CHUNK_BYTES)[pl];
      };
      const double v0 = PF(PF_VX), v1 = PF(PF_VY), v2 = PF(PF_VZ);
      double G[9];
#pragma unroll
      for (int a = 0; a < 9; a++) G[a] = PF(PF_G0 + a);
      const double b0 = ut[0][k], b1 = ut[1][k], b2 = ut[2][k];
      const double w3 = wij * p.w[k];

      // source terms, then mass matrix (adjoint_pnpn.f90:669-676)
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if (flags & FLAG_SOURCES) {
        double chi = PF(PF_RHO);
        if (flags & FLAG_RAMP) {
          if (flags & FLAG_CONVEX_UP) chi = p.f_min + (p.f_max - p.f_min) * chi * (1.0 + p.q) / (chi + p.q);
          else chi = p.f_min + (p.f_max - p.f_min) * chi / (1.0 + p.q * (1.0 - chi));
        }
        if ((flags & FLAG_CHI_OUT) && active) p.chi_out[ebase + q] = chi;
        f0 = 0.0 - v0 * chi; f1 = 0.0 - v1 * chi; f2 = 0.0 - v2 * chi;
        if (flags & FLAG_FSTATIC) { f0 += PF(PF_FS0); f1 += PF(PF_FS0 + 1); f2 += PF(PF_FS0 + 2); }
        if (flags & FLAG_LUBE) {
          const double ck = chi * p.K_lube;
          f0 += b0 * ck; f1 += b1 * ck; f2 += b2 * ck;
        }
        const double B = PF(PF_B);
        f0 *= B; f1 *= B; f2 *= B;
      } else if (flags & FLAG_FSTATIC) {
        const double B = PF(PF_B);
        f0 = PF(PF_FS0) * B; f1 = PF(PF_FS0 + 1) * B; f2 = PF(PF_FS0 + 2) * B;
      }
      if (flags & FLAG_ACCUM) { f0 += PF(PF_FIN0); f1 += PF(PF_FIN0 + 1); f2 += PF(PF_FIN0 + 2); }

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
        for (int kk = 0; kk < LX; kk++) acc[c][kk] = fma(-p.D[k + LX * kk], ft, acc[c][kk]);
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

      if (k % PC == PC - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_empty[ps]);
        if (++ps == NS) { ps = 0; pph ^= 1; }
      }
    }
    named_bar_sync(1, NCTHR);

    // ---- phase B: transposed r/s contractions of the fluxes (pencil tasks, in place) -------------
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
    named_bar_sync(1, NCTHR);

    // ---- final: f = acc - R_r - R_s, one coalesced store per component and plane ------------------
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
    named_bar_sync(1, NCTHR);
  }
}

}  // namespace b200
