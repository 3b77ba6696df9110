// Launchers of the fine-grid advection operators (advop_kernel.cuh).
#include "advop.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "advop_kernel.cuh"
#include "advop_mma_kernel.cuh"
#include "deriv_kernels.cuh"
#include "helm_kernels.cuh"

namespace b200 {
namespace {

template <int LX, int LXD, int MODE>
cudaError_t launch_one(const AdvLaunch& a) {
  using C = AdvCfg<LX, LXD, MODE>;
  static_assert(C::SMEM <= 227 * 1024, "fine-grid operator exceeds the shared memory of an SM");
  // as many CTAs per SM as shared memory allows (+1 KB reserved per CTA); cap registers to match.  The
  // register file is split over the four schedulers (16384 registers each), so what counts is the number
  // of warps per scheduler, not the CTA total (ncu r01g: 200 registers x 5 warps allowed ONE CTA per SM).
  constexpr int CTAS = (227 * 1024) / (C::SMEM + 1024) > 0 ? (227 * 1024) / (C::SMEM + 1024) : 1;
  constexpr int WPS = (C::NTHR / 32 * CTAS + 3) / 4;              // warps per scheduler
  constexpr int RCAP = 16384 / WPS / 32 / 8 * 8;
  constexpr int RMIN = (C::NTHR <= 96) ? 224 : 168;
  constexpr int MAXREG = RCAP > 255 ? 255 : (RCAP < RMIN ? RMIN : RCAP);
  AdvParams<LX, LXD> p;
  memset(&p, 0, sizeof p);
  for (int i = 0; i < LXD * LXD; i++) p.D[i] = a.D[i];
  if (LXD != LX) for (int i = 0; i < LXD * LX; i++) p.J[i] = a.J[i];
  for (int i = 0; i < LXD; i++) p.wd[i] = a.wd[i];
  for (int c = 0; c < 3; c++) { p.v[c] = a.v[c]; p.vb[c] = a.vb[c]; p.f[c] = a.f[c]; p.fs[c] = a.fs[c]; }
  for (int g = 0; g < 9; g++) p.G[g] = a.G[g];
  p.rho = a.rho; p.B = a.B; p.sens = a.sens; p.chi_out = a.chi_out;
  p.elem_list = a.elem_list; p.nelem = a.nelem; p.elem_base = a.elem_base; p.flags = a.flags;
  p.f_min = a.f_min; p.f_max = a.f_max; p.q = a.q; p.K_lube = a.K_lube; p.K_sens = a.K_sens;
  auto kern = advop_kernel<LX, LXD, MODE, MAXREG>;
  static int per_sm = 0;   // per instantiation; the function attributes are per device
  static unsigned long long dev_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !(dev_done >> dev & 1ull)) {
    if (dev < 64) dev_done |= 1ull << dev;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::NTHR, C::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  }
  const int grid = std::min(a.nelem, a.num_sm * per_sm);
  if (grid < 1) return cudaSuccess;
  kern<<<grid, C::NTHR, C::SMEM, a.stream>>>(p);
  return cudaGetLastError();
}

// lx = 8 / lxd = 12 dealiased adjoint operator on the FP64 tensor cores (advop_mma_kernel.cuh); B200_ADVOP_MMA=0
// selects the column-per-thread kernel instead (diagnostic)
bool use_mma_kernel() {
  static const int on = [] { const char* g = getenv("B200_ADVOP_MMA"); return g ? atoi(g) : 1; }();
  return on != 0;
}

cudaError_t launch_mma(const AdvLaunch& a) {
  using C = AdvMmaCfg;
  static_assert(C::SMEM <= 227 * 1024, "tensor-core dealiased operator exceeds the shared memory of an SM");
  AdvMmaParams p;
  memset(&p, 0, sizeof p);
  for (int i = 0; i < 96; i++) p.J[i] = a.J[i];
  for (int l = 0; l < 8; l++)
    for (int i = 0; i < 12; i++) {        // DJ(i,l) = sum_m D(i,m) J(m,l)
      double s = 0.0;
      for (int m = 0; m < 12; m++) s += a.D[i + 12 * m] * a.J[m + 12 * l];
      p.DJ[i + 12 * l] = s;
    }
  for (int i = 0; i < 12; i++) p.wd[i] = a.wd[i];
  for (int c = 0; c < 3; c++) { p.v[c] = a.v[c]; p.vb[c] = a.vb[c]; p.f[c] = a.f[c]; p.fs[c] = a.fs[c]; }
  for (int g = 0; g < 9; g++) p.G[g] = a.G[g];
  p.rho = a.rho; p.B = a.B; p.sens = a.sens; p.chi_out = a.chi_out;
  p.elem_list = a.elem_list; p.nelem = a.nelem; p.elem_base = a.elem_base; p.flags = a.flags;
  p.f_min = a.f_min; p.f_max = a.f_max; p.q = a.q; p.K_lube = a.K_lube; p.K_sens = a.K_sens;
  static unsigned long long dev_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !(dev_done >> dev & 1ull)) {
    if (dev < 64) dev_done |= 1ull << dev;
    cudaError_t e = cudaFuncSetAttribute(advop_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
  }
  const int grid = std::min(a.nelem, a.num_sm);
  if (grid < 1) return cudaSuccess;
  advop_mma_kernel<<<grid, C::NTHR, C::SMEM, a.stream>>>(p);
  return cudaGetLastError();
}

template <int LX>
cudaError_t launch_lx(const AdvLaunch& a, const char** msg) {
  constexpr int LXD = 3 * LX / 2;
  if (a.lxd == LX) {
    if (a.mode == ADV_LINEAR) return launch_one<LX, LX, ADV_LINEAR>(a);
    *msg = "the GLL-grid adjoint operator is the fused element kernel, not the fine-grid one";
    return cudaErrorInvalidValue;
  }
  if (a.lxd != LXD) {
    *msg = "dealiased operator: only lxd = 3*lx/2 (the factory default, advection_adjoint_fctry.f90:70) is instantiated";
    return cudaErrorInvalidValue;
  }
  if (a.mode == ADV_ADJOINT) {
    if constexpr (LX == 8) if (use_mma_kernel()) return launch_mma(a);
    return launch_one<LX, LXD, ADV_ADJOINT>(a);
  }
  return launch_one<LX, LXD, ADV_LINEAR>(a);
}

template <int LX>
cudaError_t geom_lx(int lxd, const double* J_host, const double* const src[9], double* const dst[9],
                    int nelv, int num_sm, cudaStream_t stream, const char** msg) {
  constexpr int LXD = 3 * LX / 2;
  if (lxd != LXD) {
    *msg = "dealiased operator: only lxd = 3*lx/2 is instantiated";
    return cudaErrorInvalidValue;
  }
  GeomInterpParams<LX, LXD> p;
  for (int i = 0; i < LXD * LX; i++) p.J[i] = J_host[i];
  for (int g = 0; g < 9; g++) { p.src[g] = src[g]; p.dst[g] = dst[g]; }
  p.nelv = nelv;
  constexpr int NTHR = ((LXD * LXD + 31) / 32) * 32;
  constexpr int SMEM = (LX * LX * LX + LXD * LX * LX + LXD * LXD * LX + LXD * LX) * 8;
  auto kern = geom_to_fine_kernel<LX, LXD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (e != cudaSuccess) return e;
  const long long total = 9ll * nelv;
  const int grid = (int)std::min<long long>(total, (long long)num_sm * 8);
  if (grid < 1) return cudaSuccess;
  kern<<<grid, NTHR, SMEM, stream>>>(p);
  return cudaGetLastError();
}

template <int LX>
cudaError_t deriv_lx(const DerivLaunch& a) {
  DerivParams<LX> p;
  for (int i = 0; i < LX * LX; i++) p.D[i] = a.D[i];
  for (int c = 0; c < 3; c++) { p.u[c] = a.u[c]; p.out[c] = a.out[c]; }
  for (int g = 0; g < 9; g++) p.G[g] = a.G[g];
  p.jacinv = a.jacinv; p.B = a.B; p.nelv = a.nelv;
  constexpr int NTHR = ((LX * LX + 31) / 32) * 32;
  const int grid = std::min(a.nelv, a.num_sm * 8);
  if (grid < 1) return cudaSuccess;
  if (a.mode == DERIV_CURL_B) deriv_kernel<LX, DERIV_CURL_B><<<grid, NTHR, 0, a.stream>>>(p);
  else deriv_kernel<LX, DERIV_DISSIPATION><<<grid, NTHR, 0, a.stream>>>(p);
  return cudaGetLastError();
}

template <int LX>
cudaError_t helm_lx(const HelmLaunch& a) {
  HelmParams<LX> p;
  for (int i = 0; i < LX * LX; i++) p.D[i] = a.D[i];
  for (int i = 0; i < LX; i++) p.w[i] = a.w[i];
  p.u = a.u;
  for (int g = 0; g < 9; g++) p.G[g] = a.G[g];
  p.jacinv = a.jacinv; p.B = a.B; p.out = a.out; p.h1 = a.h1; p.h2 = a.h2; p.nelv = a.nelv;
  constexpr int NTHR = ((LX * LX + 31) / 32) * 32;
  const int grid = std::min(a.nelv, a.num_sm * 6);
  if (grid < 1) return cudaSuccess;
  if (a.mode == 0) helm_kernel<LX, 0><<<grid, NTHR, 0, a.stream>>>(p);
  else helm_kernel<LX, 1><<<grid, NTHR, 0, a.stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t helm_launch(const HelmLaunch& a, const char** msg) {
  *msg = nullptr;
  switch (a.lx) {
    case 4: return helm_lx<4>(a);
    case 5: return helm_lx<5>(a);
    case 6: return helm_lx<6>(a);
    case 7: return helm_lx<7>(a);
    case 8: return helm_lx<8>(a);
    case 9: return helm_lx<9>(a);
    case 10: return helm_lx<10>(a);
    default: *msg = "lx not instantiated (4..10)"; return cudaErrorInvalidValue;
  }
}

cudaError_t deriv_launch(const DerivLaunch& a, const char** msg) {
  *msg = nullptr;
  switch (a.lx) {
    case 4: return deriv_lx<4>(a);
    case 5: return deriv_lx<5>(a);
    case 6: return deriv_lx<6>(a);
    case 7: return deriv_lx<7>(a);
    case 8: return deriv_lx<8>(a);
    case 9: return deriv_lx<9>(a);
    case 10: return deriv_lx<10>(a);
    default: *msg = "lx not instantiated (4..10)"; return cudaErrorInvalidValue;
  }
}

cudaError_t advop_launch(const AdvLaunch& a, const char** msg) {
  *msg = nullptr;
  switch (a.lx) {
    case 4: return launch_lx<4>(a, msg);
    case 5: return launch_lx<5>(a, msg);
    case 6: return launch_lx<6>(a, msg);
    case 7: return launch_lx<7>(a, msg);
    case 8: return launch_lx<8>(a, msg);
    case 9: return launch_lx<9>(a, msg);
    case 10: return launch_lx<10>(a, msg);
    default: *msg = "lx not instantiated (4..10)"; return cudaErrorInvalidValue;
  }
}

cudaError_t advop_geom_to_fine(int lx, int lxd, const double* J_host, const double* const src[9],
                               double* const dst[9], int nelv, int num_sm, cudaStream_t stream,
                               const char** msg) {
  *msg = nullptr;
  switch (lx) {
    case 4: return geom_lx<4>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 5: return geom_lx<5>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 6: return geom_lx<6>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 7: return geom_lx<7>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 8: return geom_lx<8>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 9: return geom_lx<9>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    case 10: return geom_lx<10>(lxd, J_host, src, dst, nelv, num_sm, stream, msg);
    default: *msg = "lx not instantiated (4..10)"; return cudaErrorInvalidValue;
  }
}

}  // namespace b200
