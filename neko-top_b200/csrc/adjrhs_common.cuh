// Common definitions of the fused adjoint-RHS element kernels for sm_100a (fp64): the operator they
// implement, the flag bits, and the PTX helpers (mbarrier, 1-D TMA bulk copies, L2 policies, named barriers).
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
// The three cdtp calls per component are grouped into one contravariant flux (9 instead of 27 transposed
// contractions); f and S are written once, straight from registers.
//
// Kernels: adjrhs_kernel_v3.cuh (lx = 8, DMMA fragments) and adjrhs_kernel_v2.cuh (lx = 4..7, 9, 10).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum : unsigned {
  FLAG_SOURCES = 1u,     // Brinkman term active (rho or chi present)
  FLAG_RAMP = 2u,        // the rho slot holds rho -> apply RAMP; else it already is chi
  FLAG_LUBE = 4u,        // + K_lube*chi*vb_i
  FLAG_FSTATIC = 8u,
  FLAG_ACCUM = 16u,      // f is in/out (un-fused advection_adjoint_t drop-in)
  FLAG_SENS = 32u,
  FLAG_CHI_OUT = 64u,
  FLAG_CONVEX_UP = 128u,
  FLAG_LINEAR = 256u     // (reserved)
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
// 8-byte Ampere-style asynchronous copy global -> shared (SASS: LDGSTS), completion by cp.async.wait_*
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

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

}  // namespace b200
