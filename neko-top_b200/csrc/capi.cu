// C ABI of the B200-native adjoint-RHS path (see include/neko_top_b200.h for the contract and the
// reference interfaces each entry point replaces).  No CPU fallback: every compute entry point
// launches sm_100a kernels or fails loudly.
#include "../../include/neko_top_b200.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "adjrhs_kernel_v2.cuh"
#include "adjrhs_kernel_v3.cuh"
#include "advop.h"
#include "gs_kernels.cuh"
#include "pointwise_kernels.cuh"
#include "deriv_kernels.cuh"
#include "helm_kernels.cuh"

namespace {

using namespace b200;

std::atomic<int64_t> g_launches{0};
int g_abort = 1;
thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  fprintf(stderr, "[neko_top_b200] ERROR: %s\n", buf);
  if (g_abort) abort();   // reference convention: neko_error / CUDA_CHECK end the job
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(B200_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                  cudaGetErrorString(e_));                                                        \
  } while (0)
#define NK(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      return fail(B200_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                  ncclGetErrorString(r_));                                                        \
  } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// NVTX ranges named after the reference's profiler regions (Neko's profiler_start_region is nvtxRangePush on
// the CUDA backend): 'Fluid' (adjoint/adjoint_pnpn.f90:639) holds the RHS construction, 'Velocity residual'
// (:746) the gs_op at :755-757.  Header-only NVTX3: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

constexpr int LX_MAX = 10;

struct Handle {
  int lx = 0, nelv = 0, device = 0;
  int64_t n = 0;
  int num_sm = 0;
  cudaStream_t stream = nullptr;
  bool have_space = false, have_geom = false;
  double D[LX_MAX * LX_MAX];
  double w[LX_MAX];
  const double* G[9] = {};
  const double* B = nullptr;
  double* geom_pack = nullptr;    // private image [e][k][10][lx*lx] of G[0..8], B (adjrhs_kernel_v2.cuh)
  double f_min = 0.0, f_max = 1000.0, q = 1.0, K_lube = 1.0, K_sens = 1.0;
  int convex_up = 1, if_lube = 1;
  const int* lube_mask = nullptr;
  int lube_mask_size = 0;
  int cfg = -1;   // kernel configuration override (B200_ADJRHS_CFG)

  // dealiased operator: the state of adv_lin_dealias_t (adjoint/adv_adjoint_dealias.f90:56-131)
  int lxd = 0;                     // 0: b200_adv_dealias_init not called
  bool dealias_fused = false;      // b200_adjrhs_compute/step use the dealiased operator
  std::vector<double> Jd, Dd, wd;  // GLL_to_GL matrix (lxd x lx), Xh_GL%dx, Xh_GL%wx
  double* G_fine = nullptr;        // coef_GL: 9 arrays of nelv*lxd^3 (one allocation)

  // gather-scatter
  bool have_gs = false;
  int nclass = 0;
  int64_t nmember = 0;
  int *gs_off = nullptr, *gs_dof = nullptr, *gs_rep = nullptr;
  unsigned char* gs_skip = nullptr;     // classes handled by the shared-node path
  int* gs_shared_cls = nullptr;         // list of local classes that contain a shared node
  int n_shared_cls = 0;

  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_recv = nullptr;
  int nshared = 0, nneigh = 0;
  std::vector<int> neigh_rank, neigh_off;
  int *d_send_dof = nullptr, *d_shared_dof = nullptr, *d_s_class = nullptr, *d_c_off = nullptr,
      *d_c_src = nullptr;
  double *d_send = nullptr, *d_recv = nullptr;
  int nsend = 0;
  int *d_bnd_elem = nullptr, *d_int_elem = nullptr;
  int nbnd = 0, nint = 0;

  // class schedule of the pipelined host step / the packed-list pass: element processing order and the
  // class schedule built for it by build_gs_schedule()
  std::vector<int> order;          // processing order of the elements (empty: 0..nelv-1)
  int* d_order = nullptr;
  int gs_mode = 0;                 // 0: CSR class lists (default), 1: class lists packed by size (B200_GS_MODE=1)
  bool overlap_elem = false;       // B200_EXCHANGE_OVERLAP=elem: overlap the exchange with the interior-element kernel
  bool exchange_side = false;      // B200_EXCHANGE_SIDE=1: shared-class sum + pack on the communication stream (r02n: no gain)
  int gs_un = 1;                   // classes in flight per thread in gs_op_kernel (B200_GS_UN: 1, 2, 4; measured: 1 is best)
  bool sched_valid = false;
  int sched_nelem = 0;             // length of the element list the schedule was built for
  int sched_kind = 0;              // 1: all elements (single GPU), 2: interior elements (multi-GPU split)
  int *sched_eoff = nullptr, *sched_pair = nullptr, *sched_quad = nullptr, *sched_oct = nullptr,
      *sched_hex = nullptr, *sched_left = nullptr;
  int sched_nleft = 0;
  int sched_n2 = 0, sched_n4 = 0, sched_n8 = 0, sched_n16 = 0;
  int64_t sched_nfused = 0;        // classes in the packed lists (2..16 members)
  std::vector<int> sched_elem_last;   // host copy: per position, the last position whose classes touch it (kind 1)

  // x stage of the lx = 8 element kernel (adjrhs_kernel_v3.cuh XS): per-element link flags and the CSR lists
  // of the classes that stay in the gather-scatter pass
  // staged direct-stiffness summation (gs_kernels.cuh "staged"): 0 off, 1 = x pairs inside the lx = 8 element
  // kernel (bit-identical to the plain pass), 2 = x in the kernel + y and z face passes (product classes)
  int xs_enable = 2;               // B200_XSTAGE
  bool xs_valid = false;
  int xs_level = 0;                // level the lists below were built for
  int xs_nslots = 0;               // element slots of the grid the links were built for
  int face_ctas = 8;               // CTAs of 256 threads per SM of the y / z face passes (B200_FACE_CTAS)
  int xs_nolink = 0;               // B200_XS_NOLINK=1 (diagnostic): same element map, no class staged
  unsigned long long* xs_mask[3] = {};   // per element: nodes of the i/j/k = 0 face summed with the neighbour's 7 face
  int* xs_pred[3] = {};            // neighbour element across that face
  bool xs_have[3] = {};            // any bit set in direction a
  int *xs_off = nullptr, *xs_dof = nullptr;
  unsigned char* xs_skip = nullptr;
  int xs_nclass = 0;
  int64_t xs_nlinked = 0, xs_nmember = 0;

  // minimum-dissipation objective chain: six work fields (curl of curl) and reduction partials
  double* work6 = nullptr;
  double* d_partial = nullptr;

  // host-staged step
  double* stage[11] = {};
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> ev_h2d, ev_k;
  cudaEvent_t ev_done = nullptr;

  // timing
  bool phase_timing = false;      // B200_PHASE_TIMING=1: events between the phases of the last step
  cudaEvent_t pev[10] = {};
  int pev_n = 0;
  bool timing = false;
  std::vector<cudaEvent_t> tev;   // pool of (start, mid, end) triples: elem kernel = mid-start, gs = end-mid
  size_t tev_n = 0;               // events of the pool recorded since timing was enabled
};

Handle* H(void* h) { return reinterpret_cast<Handle*>(h); }

int grid_for(int64_t n, int threads, int num_sm, int per_sm) {
  int64_t need = (n + threads - 1) / threads;
  int64_t cap = (int64_t)num_sm * per_sm;
  return (int)std::max<int64_t>(1, std::min(need, cap));
}

// ---------------------------------------------------------------------------------------------
// fused element kernel launcher
// ---------------------------------------------------------------------------------------------
struct LaunchArgs {
  const double* v[3];
  const double* vb[3];
  const double* rho;      // rho or chi
  bool rho_is_chi;
  const double* fs[3];
  const double* fin[3];   // accumulate mode
  double* f[3];
  double* sens;
  double* chi_out;
  const int* elem_list;
  int nelem;
  int elem_begin;         // used only when elem_list == nullptr (via pointer offsets)
  bool sources;
  bool no_dealias;         // un-fused GLL-grid drop-in: ignore the handle's dealias switch
  bool xstage;             // v3 XS: contiguous runs per slot, i-face pair classes summed in the kernel
};

// ---- second-generation kernel (adjrhs_kernel_v2.cuh): NE element slots + one TMA warp per SM ----------
template <int LX, int NE, int NS, int NF, int MAXREG, int NPROD = 1>
int launch_v2_cfg(Handle* h, const LaunchArgs& a) {
  using C = V2Cfg<LX, NE, NS, NF, NPROD>;
  static_assert(C::SMEM <= 227 * 1024, "v2 configuration exceeds the shared memory of an SM");
  static_assert(C::NTHREADS <= 1024, "v2 configuration exceeds 1024 threads");
  KParams2<LX> p;
  memset(&p, 0, sizeof p);
  for (int i = 0; i < LX * LX; i++) p.D[i] = h->D[i];
  for (int i = 0; i < LX; i++) p.w[i] = h->w[i];
  const size_t eoff = 0;             // the element offset goes in as an index (alignment for odd lx)
  for (int c = 0; c < 3; c++) p.ub[c] = a.vb[c] + eoff;
  unsigned flags = 0;
  if (!h->geom_pack) return fail(B200_ERR_STATE, "packed geometry image missing (set_geometry)");
  p.geom = h->geom_pack + eoff * NGEO;
  for (int c = 0; c < 3; c++) p.pf[R_V - NGEO + c] = a.v[c] + eoff;
  if (a.sources) {
    p.pf[R_RHO - NGEO] = a.rho + eoff;
    flags |= FLAG_SOURCES;
    if (!a.rho_is_chi) flags |= FLAG_RAMP;
    if (h->convex_up) flags |= FLAG_CONVEX_UP;
    if (h->if_lube && h->lube_mask_size == 0) flags |= FLAG_LUBE;
    if (a.chi_out) flags |= FLAG_CHI_OUT;
  }
  if constexpr (NF >= NF_FULL) {
    if (a.fs[0]) {
      for (int c = 0; c < 3; c++) p.pf[R_FS - NGEO + c] = a.fs[c] + eoff;
      flags |= FLAG_FSTATIC;
    }
    if (a.fin[0]) {
      for (int c = 0; c < 3; c++) p.pf[R_FIN - NGEO + c] = a.fin[c] + eoff;
      flags |= FLAG_ACCUM;
    }
  } else if (a.fs[0] || a.fin[0]) {
    return fail(B200_ERR_STATE, "internal: static forcing / accumulate mode need the NF_FULL kernel");
  }
  if (a.sens) flags |= FLAG_SENS;
  int na = NGEO;
  for (int i = 0; i < NF_FULL - NGEO; i++) na += (p.pf[i] != nullptr);
  p.n_active = na;
  for (int c = 0; c < 3; c++) p.f[c] = a.f[c] + eoff;
  p.sens = a.sens ? a.sens + eoff : nullptr;
  p.chi_out = a.chi_out ? a.chi_out + eoff : nullptr;
  p.elem_list = a.elem_list;
  p.nelem = a.nelem;
  p.elem_base = a.elem_begin;
  p.flags = flags;
  p.f_min = h->f_min; p.f_max = h->f_max; p.q = h->q; p.K_lube = h->K_lube;
  p.K_sens = h->if_lube ? h->K_sens : 0.0;

  auto kern = adjrhs_v2_kernel<LX, NE, NS, NF, MAXREG, NPROD>;
  static unsigned long long attr_set = 0;   // per instantiation, one bit per device (the attribute is per device)
  if (h->device >= 64 || !(attr_set >> h->device & 1ull)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    if (h->device < 64) attr_set |= 1ull << h->device;
  }
  const int grid = std::min((a.nelem + NE - 1) / NE, h->num_sm);
  if (grid < 1) return B200_OK;
  kern<<<grid, C::NTHREADS, C::SMEM, h->stream>>>(p);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

template <int LX, int NE, int NS, int MAXREG, int NE_FULL = NE, int NS_FULL = NS, int NPROD = 1>
int launch_v2(Handle* h, const LaunchArgs& a) {
  if (a.fs[0] || a.fin[0]) return launch_v2_cfg<LX, NE_FULL, NS_FULL, NF_FULL, MAXREG, NPROD>(h, a);
  return launch_v2_cfg<LX, NE, NS, NF_FUSED, MAXREG, NPROD>(h, a);
}

// ---- third-generation kernel (adjrhs_kernel_v3.cuh): lx = 8, DMMA contractions --------------------------
template <int NF>
int fill_params2_lx8(Handle* h, const LaunchArgs& a, KParams2<8>& p) {
  memset(&p, 0, sizeof p);
  for (int i = 0; i < 64; i++) p.D[i] = h->D[i];
  for (int i = 0; i < 8; i++) p.w[i] = h->w[i];
  const size_t eoff = (size_t)a.elem_begin * 512;
  for (int c = 0; c < 3; c++) p.ub[c] = a.vb[c] + eoff;
  unsigned flags = 0;
  if (!h->geom_pack) return fail(B200_ERR_STATE, "packed geometry image missing (set_geometry)");
  p.geom = h->geom_pack + eoff * NGEO;
  for (int c = 0; c < 3; c++) p.pf[R_V - NGEO + c] = a.v[c] + eoff;
  if (a.sources) {
    p.pf[R_RHO - NGEO] = a.rho + eoff;
    flags |= FLAG_SOURCES;
    if (!a.rho_is_chi) flags |= FLAG_RAMP;
    if (h->convex_up) flags |= FLAG_CONVEX_UP;
    if (h->if_lube && h->lube_mask_size == 0) flags |= FLAG_LUBE;
    if (a.chi_out) flags |= FLAG_CHI_OUT;
  }
  if constexpr (NF >= NF_FULL) {
    if (a.fs[0]) {
      for (int c = 0; c < 3; c++) p.pf[R_FS - NGEO + c] = a.fs[c] + eoff;
      flags |= FLAG_FSTATIC;
    }
    if (a.fin[0]) {
      for (int c = 0; c < 3; c++) p.pf[R_FIN - NGEO + c] = a.fin[c] + eoff;
      flags |= FLAG_ACCUM;
    }
  } else if (a.fs[0] || a.fin[0]) {
    return fail(B200_ERR_STATE, "internal: static forcing / accumulate mode need the NF_FULL kernel");
  }
  if (a.sens) flags |= FLAG_SENS;
  int na = NGEO;
  for (int i = 0; i < NF_FULL - NGEO; i++) na += (p.pf[i] != nullptr);
  p.n_active = na;
  for (int c = 0; c < 3; c++) p.f[c] = a.f[c] + eoff;
  p.sens = a.sens ? a.sens + eoff : nullptr;
  p.chi_out = a.chi_out ? a.chi_out + eoff : nullptr;
  p.elem_list = a.elem_list;
  p.nelem = a.nelem;
  p.flags = flags;
  p.f_min = h->f_min; p.f_max = h->f_max; p.q = h->q; p.K_lube = h->K_lube;
  p.K_sens = h->if_lube ? h->K_sens : 0.0;
  if (a.xstage) {
    if (!h->xs_valid || a.elem_begin != 0 || a.nelem != h->nelv || a.elem_list)
      return fail(B200_ERR_STATE, "internal: x stage without matching link flags");
    p.xmask = h->xs_mask[0];
  }
  return B200_OK;
}

constexpr int XS_NE = 3;   // element slots per CTA of the x-stage kernels (the default configuration)
template <int NE, int NW, int DS, int NF, int MAXREG, bool LIST, int XS = 0>
int launch_v3_cfg2(Handle* h, const LaunchArgs& a) {
  using C = V3Cfg<NE, NW, DS, NF>;
  if (a.xstage != (XS != 0)) return fail(B200_ERR_STATE, "internal: v3 kernel variant / x stage mismatch");
  constexpr int SMEM = C::SMEM;
  static_assert(SMEM <= 227 * 1024, "v3 configuration exceeds the shared memory of an SM");
  static_assert(C::NTHREADS <= 1024, "v3 configuration exceeds 1024 threads");
  static_assert(C::NTHREADS * MAXREG <= 65536, "v3 configuration exceeds the register file");
  KParams2<8> p;
  if (int r = fill_params2_lx8<NF>(h, a, p)) return r;
  auto kern = adjrhs_v3_kernel<NE, NW, DS, NF, MAXREG, LIST, XS>;
  static unsigned long long attr_set = 0;   // per instantiation, one bit per device (the attribute is per device)
  if (h->device >= 64 || !(attr_set >> h->device & 1ull)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (h->device < 64) attr_set |= 1ull << h->device;
  }
  const int grid = std::min((a.nelem + NE - 1) / NE, h->num_sm);
  if (grid < 1) return B200_OK;
  if (XS && grid * NE != h->xs_nslots) return fail(B200_ERR_STATE, "internal: x-stage links built for another grid");
  kern<<<grid, C::NTHREADS, SMEM, h->stream>>>(p);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

template <int NE, int NW, int DS, int NF, int MAXREG>
int launch_v3_cfg(Handle* h, const LaunchArgs& a) {
  if (a.elem_list) return launch_v3_cfg2<NE, NW, DS, NF, MAXREG, true>(h, a);
  return launch_v3_cfg2<NE, NW, DS, NF, MAXREG, false>(h, a);
}

template <int NE, int NW, int DS, int MAXREG, int NE_FULL = NE, int DS_FULL = DS>
int launch_v3(Handle* h, const LaunchArgs& a) {
  if (a.fs[0] || a.fin[0]) return launch_v3_cfg<NE_FULL, NW, DS_FULL, NF_FULL, MAXREG>(h, a);
  return launch_v3_cfg<NE, NW, DS, NF_FUSED, MAXREG>(h, a);
}
// the default configuration (with or without the x stage)
int launch_v3_default(Handle* h, const LaunchArgs& a) {
  const bool full = a.fs[0] || a.fin[0];
  if (a.xstage) {
    if (a.elem_list) return fail(B200_ERR_STATE, "internal: x stage with an element list");
    if (full) return launch_v3_cfg2<XS_NE, 4, 1, NF_FULL, 168, false, 2>(h, a);
    return launch_v3_cfg2<XS_NE, 4, 2, NF_FUSED, 168, false, 2>(h, a);
  }
  if (full) return launch_v3_cfg<3, 4, 1, NF_FULL, 168>(h, a);
  return launch_v3_cfg<3, 4, 2, NF_FUSED, 168>(h, a);
}

// fine-grid operators (advop_kernel.cuh): dealiased adjoint / linearised advection, GLL-grid linearised
int launch_fine(Handle* h, const LaunchArgs& a, int mode, bool dealias) {
  if (!h->have_space || !h->have_geom) return fail(B200_ERR_STATE, "set_space/set_geometry not called");
  if (dealias && h->lxd == 0) return fail(B200_ERR_STATE, "b200_adv_dealias_init not called");
  AdvLaunch L;
  memset(&L, 0, sizeof L);
  L.lx = h->lx; L.mode = mode;
  const size_t nd = dealias ? (size_t)h->nelv * h->lxd * h->lxd * h->lxd : (size_t)h->n;
  if (dealias) {
    L.lxd = h->lxd; L.D = h->Dd.data(); L.J = h->Jd.data(); L.wd = h->wd.data();
    for (int g = 0; g < 9; g++) L.G[g] = h->G_fine + g * nd;
  } else {
    L.lxd = h->lx; L.D = h->D; L.J = nullptr; L.wd = h->w;
    for (int g = 0; g < 9; g++) L.G[g] = h->G[g];
  }
  unsigned flags = 0;
  for (int c = 0; c < 3; c++) { L.v[c] = a.v[c]; L.vb[c] = a.vb[c]; L.f[c] = a.f[c]; L.fs[c] = a.fs[c]; }
  if (a.fin[0]) {
    flags |= FLAG_ACCUM;     // advection_adjoint_t contract: f in/out
  } else {
    if (a.sources) {
      L.rho = a.rho;
      flags |= FLAG_SOURCES;
      if (!a.rho_is_chi) flags |= FLAG_RAMP;
      if (h->convex_up) flags |= FLAG_CONVEX_UP;
      if (h->if_lube && h->lube_mask_size == 0) flags |= FLAG_LUBE;
      if (a.chi_out) flags |= FLAG_CHI_OUT;
    }
    if (a.fs[0]) flags |= FLAG_FSTATIC;
    if (a.sens) flags |= FLAG_SENS;
  }
  L.B = h->B; L.sens = a.sens; L.chi_out = a.chi_out;
  L.elem_list = a.elem_list; L.nelem = a.nelem; L.elem_base = a.elem_list ? 0 : a.elem_begin; L.flags = flags;
  L.f_min = h->f_min; L.f_max = h->f_max; L.q = h->q; L.K_lube = h->K_lube;
  L.K_sens = h->if_lube ? h->K_sens : 0.0;
  L.num_sm = h->num_sm; L.stream = h->stream;
  const char* msg = nullptr;
  cudaError_t e = advop_launch(L, &msg);
  if (e != cudaSuccess)
    return fail(msg ? B200_ERR_ARG : B200_ERR_CUDA, "fine-grid advection operator (lx=%d lxd=%d mode=%d): %s",
                L.lx, L.lxd, mode, msg ? msg : cudaGetErrorString(e));
  if (a.nelem > 0) LAUNCHED();
  return B200_OK;
}

int launch_fused(Handle* h, const LaunchArgs& a) {
  if (!h->have_space || !h->have_geom) return fail(B200_ERR_STATE, "set_space/set_geometry not called");
  if (h->dealias_fused && !a.no_dealias) return launch_fine(h, a, ADV_ADJOINT, true);
  const int cfg = h->cfg;
  switch (h->lx) {
    case 4: return launch_v2<4, 8, 4, 224>(h, a);
    // lx != 8 (v2): configurations from the r01j sweep (tools/lxsweep.py, profiles/r01j_lxsweep.jsonl).  What
    // matters most is a warp count whose per-scheduler register share avoids spills (the register file is
    // split over four schedulers: 9-12 warps -> 168 registers, <= 8 warps -> 255); cfg 32 = the older choice.
    // (deeper plane rings -- 3..7 stages -- were measured in r02j: no gain at lx = 5, 6, 9, 10, +6 % at lx = 7 with 7
    // stages; the consumers' waits on the ring are the single producer lane's issue rate, not the ring depth)
    // lx = 5, 6: several producer warps (ncu r02h / sweeps r02n, r02o): one elected lane issuing ~5 bulk copies per
    // plane and slot could not keep the slots fed -- the consumers spun on the ring's mbarriers.  lx = 5: 8 slots
    // + 4 producer warps 10.3 -> 16.9 GDOF/s; lx = 6: 4 slots + 4 producer warps 12.0 -> 14.8 GDOF/s.
    // (2, 3, 5, 6 producer warps and other slot counts: r02n-r02p, all slower; cfg 32 = the round-1 single-producer choice)
    case 5: return cfg == 32 ? launch_v2<5, 11, 2, 168>(h, a) : launch_v2<5, 8, 3, 168, 8, 2, 4>(h, a);
    case 6: return cfg == 32 ? launch_v2<6, 5, 2, 168>(h, a) : launch_v2<6, 4, 3, 168, 4, 2, 4>(h, a);
    case 7: return cfg == 32 ? launch_v2<7, 4, 2, 224>(h, a) : launch_v2<7, 3, 7, 255, 3, 6>(h, a);
    case 8:
      switch (cfg) {
        case 3: return launch_v2<8, 3, 4, 255, 3, 4>(h, a);   // the DFMA (v2) kernel at lx = 8, for A/B runs
        case 20: return launch_v3<4, 4, 1, 128, 3, 1>(h, a); // 16 warps (A/B runs, profiles r01c)
        case 25: return launch_v3<3, 4, 1, 168>(h, a);       // 12 warps, one plane stage
        default: return launch_v3_default(h, a);
      }
    // (two producer warps at lx = 7, 9, 10: -3 % / 0 / +3 %, r02n: these orders are bound by the consumers)
    case 9: return cfg == 32 ? launch_v2<9, 3, 2, 200>(h, a) : launch_v2<9, 2, 4, 255, 2, 2>(h, a);
    case 10: return launch_v2<10, 2, 2, 224>(h, a);
    default: return fail(B200_ERR_ARG, "lx=%d not instantiated (4..10)", h->lx);
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_fields(std::initializer_list<const void*> ps) {
  for (const void* p : ps)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15) != 0)
      return fail(B200_ERR_ARG, "device field pointer %p is not 16-byte aligned", p);
  return B200_OK;
}

int time_mark(Handle* h) {
  if (!h->timing) return B200_OK;
  if (h->tev_n >= 3 * 4096) return B200_OK;
  if (h->tev_n == h->tev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->tev.push_back(e);
  }
  CK(cudaEventRecord(h->tev[h->tev_n++], h->stream));
  return B200_OK;
}

// diagnostic: event after phase `idx` of the current step (multi-GPU split path)
int phase_mark(Handle* h, int idx) {
  if (!h->phase_timing || idx >= 10) return B200_OK;
  if (!h->pev[idx]) CK(cudaEventCreate(&h->pev[idx]));
  CK(cudaEventRecord(h->pev[idx], h->stream));
  h->pev_n = idx + 1;
  return B200_OK;
}

// local gather-scatter launches
int gs_launch(Handle* h, double* f0, double* f1, double* f2, int nf) {
  if (!h->have_gs) return fail(B200_ERR_STATE, "b200_gs_init not called");
  if (h->nclass == 0) return B200_OK;
  const int threads = 256;
  const int grid = grid_for(h->nclass, threads, h->num_sm, 8);
  // 3 CTAs of 256 threads per SM, one grid-stride pass: measured best (r01n: 1.60 ms at 64^3 against 1.68 with 8
  // CTAs/SM and 1.81 with 2) -- the pass is bound by memory requests in flight, and more resident warps only
  // spread them over more DRAM pages
  const int grid3 = grid_for(h->nclass, threads, h->num_sm, 3);
  if (nf == 1) gs_op_kernel<1, 1, 3><<<grid3, threads, 0, h->stream>>>(f0, f0, f0, h->gs_off, h->gs_dof, h->nclass);
  else if (h->gs_un == 2) gs_op_kernel<3, 2, 2><<<grid_for(h->nclass, threads, h->num_sm, 2), threads, 0, h->stream>>>(
      f0, f1, f2, h->gs_off, h->gs_dof, h->nclass);
  else if (h->gs_un == 4) gs_op_kernel<3, 4, 2><<<grid_for(h->nclass, threads, h->num_sm, 2), threads, 0, h->stream>>>(
      f0, f1, f2, h->gs_off, h->gs_dof, h->nclass);
  else gs_op_kernel<3, 1, 3><<<grid3, threads, 0, h->stream>>>(f0, f1, f2, h->gs_off, h->gs_dof, h->nclass);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

// rank-local pass of the fused step: all classes (minus the shared-node classes of a multi-GPU run), or, after
// an element kernel with the x stage, the classes that kernel left over
int gs_step_pass(Handle* h, double* f0, double* f1, double* f2, bool xs) {
  const int nc = xs ? h->xs_nclass : h->nclass;
  if (nc == 0) return B200_OK;
  const int threads = 256, grid = grid_for(nc, threads, h->num_sm, 3);
  if (xs && h->gs_un == 2)
    gs_op_kernel<3, 2, 2><<<grid_for(nc, threads, h->num_sm, 2), threads, 0, h->stream>>>(f0, f1, f2, h->xs_off, h->xs_dof, nc, h->xs_skip);
  else if (xs && h->gs_un == 4)
    gs_op_kernel<3, 4, 2><<<grid_for(nc, threads, h->num_sm, 2), threads, 0, h->stream>>>(f0, f1, f2, h->xs_off, h->xs_dof, nc, h->xs_skip);
  else if (xs && h->gs_un == 8)
    gs_op_kernel<3, 1, 3><<<grid_for(nc, threads, h->num_sm, 8), threads, 0, h->stream>>>(f0, f1, f2, h->xs_off, h->xs_dof, nc, h->xs_skip);
  else if (xs)
    gs_op_kernel<3, 1, 3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->xs_off, h->xs_dof, nc, h->xs_skip);
  else
    gs_op_kernel<3, 1, 3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->gs_off, h->gs_dof, nc, h->gs_skip);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

// shared-node exchange: pack -> ncclSend/Recv (comm stream) -> unpack
// shared_first: first sum the classes that hold shared nodes; side_stream: run that sum and the pack on the
// communication stream too (behind an event on the main stream), so that the main stream goes straight on to
// the rank-local passes -- they touch disjoint nodes
int gs_exchange(Handle* h, double* f0, double* f1, double* f2, int nf, bool shared_first = false,
                bool side_stream = false) {
  if (!h->comm || h->nshared == 0) return B200_OK;
  const int threads = 256;
  cudaStream_t ps = side_stream ? h->comm_stream : h->stream;
  if (side_stream) {
    CK(cudaEventRecord(h->ev_pack, h->stream));
    CK(cudaStreamWaitEvent(h->comm_stream, h->ev_pack, 0));
  }
  if (shared_first && h->n_shared_cls > 0) {
    const int gl = grid_for(h->n_shared_cls, threads, h->num_sm, 4);
    gs_op_list_kernel<3><<<gl, threads, 0, ps>>>(f0, f1, f2, h->gs_off, h->gs_dof, h->gs_shared_cls, h->n_shared_cls);
    LAUNCHED();
  }
  const int grid = grid_for(h->nsend, threads, h->num_sm, 4);
  if (nf == 1) gs_pack_kernel<1><<<grid, threads, 0, ps>>>(f0, f0, f0, h->d_send_dof, h->nsend, h->d_send);
  else gs_pack_kernel<3><<<grid, threads, 0, ps>>>(f0, f1, f2, h->d_send_dof, h->nsend, h->d_send);
  LAUNCHED();
  CK(cudaGetLastError());
  if (!side_stream) {
    CK(cudaEventRecord(h->ev_pack, h->stream));
    CK(cudaStreamWaitEvent(h->comm_stream, h->ev_pack, 0));
  }
  NK(ncclGroupStart());
  for (int j = 0; j < h->nneigh; j++) {
    const size_t o = (size_t)h->neigh_off[j] * nf, cnt = (size_t)(h->neigh_off[j + 1] - h->neigh_off[j]) * nf;
    NK(ncclSend(h->d_send + o, cnt, ncclDouble, h->neigh_rank[j], h->comm, h->comm_stream));
    NK(ncclRecv(h->d_recv + o, cnt, ncclDouble, h->neigh_rank[j], h->comm, h->comm_stream));
  }
  NK(ncclGroupEnd());
  CK(cudaEventRecord(h->ev_recv, h->comm_stream));
  return B200_OK;
}
int gs_finish_exchange(Handle* h, double* f0, double* f1, double* f2, int nf) {
  if (!h->comm || h->nshared == 0) return B200_OK;
  CK(cudaStreamWaitEvent(h->stream, h->ev_recv, 0));
  const int threads = 256;
  const int grid = grid_for(h->nshared, threads, h->num_sm, 4);
  if (nf == 1)
    gs_unpack_kernel<1><<<grid, threads, 0, h->stream>>>(f0, f0, f0, h->d_recv, h->d_c_off, h->d_c_src,
                                                          h->d_shared_dof, h->d_s_class, h->gs_off,
                                                          h->gs_dof, h->nshared);
  else
    gs_unpack_kernel<3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->d_recv, h->d_c_off, h->d_c_src,
                                                          h->d_shared_dof, h->d_s_class, h->gs_off,
                                                          h->gs_dof, h->nshared);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

template <typename T>
int dmalloc(T** p, size_t count) {
  CK(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T)));
  return B200_OK;
}

// frees device temporaries on every exit path of a set-up routine
struct DevTemps {
  std::vector<void*> ptrs;
  ~DevTemps() { for (void* q : ptrs) cudaFree(q); }
  template <typename T>
  int alloc(T** q, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(q), std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "cudaMalloc of a set-up buffer failed: %s", cudaGetErrorString(e));
    ptrs.push_back(*q);
    return B200_OK;
  }
  // free one buffer early (peak memory of the set-up) and forget it
  template <typename T>
  void release(T*& q) {
    for (auto& x : ptrs) if (x == (void*)q) { cudaFree(x); x = nullptr; }
    q = nullptr;
  }
};

// true if launch_fused() will run the default v3 element kernel (the one with an x-stage variant)
bool uses_v3(const Handle* h) {
  return h->lx == 8 && !h->dealias_fused && (h->cfg <= 0 || (h->cfg > 25 && h->cfg < 100));   // default config
}

void free_schedule(Handle* h) {
  cudaFree(h->sched_eoff); cudaFree(h->sched_pair); cudaFree(h->sched_quad); cudaFree(h->sched_oct);
  cudaFree(h->sched_hex); cudaFree(h->sched_left);
  h->sched_eoff = h->sched_pair = h->sched_quad = h->sched_oct = h->sched_hex = h->sched_left = nullptr;
  h->sched_valid = false; h->sched_nleft = 0; h->sched_nfused = 0; h->sched_nelem = 0; h->sched_kind = 0;
}

// Class schedule (classes sorted by the position of the element that completes them, packed by size) for the element list (list, nlist):
// what the pipelined host step and the packed-list pass walk
// (list == nullptr: elements 0..nlist-1 in order).  Elements outside the list count as already stored.
int build_gs_schedule(Handle* h, const int* list, int nlist) {
  free_schedule(h);
  if (!h->have_gs || h->nclass == 0 || nlist == 0) return B200_OK;
  cudaStream_t st = h->stream;
  const int nc = h->nclass, threads = 256;
  int *d_pos = nullptr, *d_cls = nullptr, *d_cls2 = nullptr, *d_bstart = nullptr;
  unsigned long long *d_key = nullptr, *d_key2 = nullptr;
  if (int r = dmalloc(&d_pos, (size_t)h->nelv)) return r;
  if (int r = dmalloc(&d_cls, (size_t)nc)) return r;
  if (int r = dmalloc(&d_cls2, (size_t)nc)) return r;
  if (int r = dmalloc(&d_key, (size_t)nc)) return r;
  if (int r = dmalloc(&d_key2, (size_t)nc)) return r;
  if (int r = dmalloc(&d_bstart, 8)) return r;
  CK(cudaMemsetAsync(d_pos, 0xff, sizeof(int) * (size_t)h->nelv, st));
  gs_pos_kernel<<<grid_for(nlist, threads, h->num_sm, 8), threads, 0, st>>>(list, nlist, d_pos);
  LAUNCHED();
  gs_class_key_kernel<<<grid_for(nc, threads, h->num_sm, 8), threads, 0, st>>>(
      h->gs_off, h->gs_dof, h->gs_skip, nc, d_pos, h->lx * h->lx * h->lx, d_key, d_cls);
  LAUNCHED();
  CK(cudaGetLastError());
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_cls, d_cls2, nc, 0, 44, st));
  CK(cudaMalloc(&d_tmp, tmp_bytes));
  CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_cls, d_cls2, nc, 0, 44, st));
  CK(cudaFree(d_tmp));
  if (int r = dmalloc(&h->sched_eoff, 4 * ((size_t)nlist + 1))) return r;
  gs_eoff_kernel<<<grid_for(4ll * (nlist + 1) + 7, threads, h->num_sm, 8), threads, 0, st>>>(
      d_key2, nc, nlist, h->sched_eoff, d_bstart);
  LAUNCHED();
  CK(cudaGetLastError());
  int bstart[8] = {};
  CK(cudaMemcpyAsync(bstart, d_bstart, sizeof(int) * 7, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const size_t n2 = bstart[1] - bstart[0], n4 = bstart[2] - bstart[1], n8 = bstart[3] - bstart[2],
               n16 = bstart[4] - bstart[3], nl = bstart[5] - bstart[4];
  if (int r = dmalloc(&h->sched_pair, 2 * n2 + 4)) return r;
  if (int r = dmalloc(&h->sched_quad, 4 * n4 + 4)) return r;
  if (int r = dmalloc(&h->sched_oct, 8 * n8 + 4)) return r;
  if (int r = dmalloc(&h->sched_hex, 16 * n16 + 4)) return r;
  if (int r = dmalloc(&h->sched_left, nl + 1)) return r;
  if (bstart[5] > 0) {
    gs_fill_kernel<<<grid_for(bstart[5], threads, h->num_sm, 8), threads, 0, st>>>(
        d_cls2, d_bstart, h->gs_off, h->gs_dof, h->sched_pair, h->sched_quad, h->sched_oct, h->sched_hex,
        h->sched_left);
    LAUNCHED();
    CK(cudaGetLastError());
  }
  h->sched_elem_last.clear();
  if (!list && nlist == h->nelv) {          // mesh order: what the pipelined host step needs
    int* d_last = nullptr;
    if (int r = dmalloc(&d_last, (size_t)h->nelv)) return r;
    gs_iota_kernel<<<grid_for(h->nelv, threads, h->num_sm, 8), threads, 0, st>>>(d_last, h->nelv);
    LAUNCHED();
    gs_elem_last_kernel<<<grid_for(nc, threads, h->num_sm, 8), threads, 0, st>>>(
        h->gs_off, h->gs_dof, h->gs_skip, nc, d_pos, h->lx * h->lx * h->lx, h->nelv - 1, d_last);
    LAUNCHED();
    CK(cudaGetLastError());
    h->sched_elem_last.resize(h->nelv);
    CK(cudaMemcpyAsync(h->sched_elem_last.data(), d_last, sizeof(int) * (size_t)h->nelv, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_last));
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(d_pos)); CK(cudaFree(d_cls)); CK(cudaFree(d_cls2)); CK(cudaFree(d_key)); CK(cudaFree(d_key2));
  CK(cudaFree(d_bstart));
  h->sched_nleft = (int)nl;
  h->sched_n2 = (int)n2; h->sched_n4 = (int)n4; h->sched_n8 = (int)n8; h->sched_n16 = (int)n16;
  h->sched_nfused = (int64_t)(n2 + n4 + n8 + n16);
  h->sched_nelem = nlist;
  h->sched_valid = true;
  return B200_OK;
}


// multi-GPU shared-node state (b200_gs_init_shared); it refers to the class list of the last b200_gs_init
void free_shared(Handle* h) {
  cudaFree(h->d_send_dof); cudaFree(h->d_shared_dof); cudaFree(h->d_s_class); cudaFree(h->d_c_off);
  cudaFree(h->d_c_src); cudaFree(h->d_send); cudaFree(h->d_recv); cudaFree(h->gs_skip);
  cudaFree(h->gs_shared_cls);
  h->d_send_dof = h->d_shared_dof = h->d_s_class = h->d_c_off = h->d_c_src = nullptr;
  h->d_send = h->d_recv = nullptr; h->gs_skip = nullptr; h->gs_shared_cls = nullptr;
  h->n_shared_cls = 0; h->nsend = 0; h->nshared = 0; h->nneigh = 0;
  h->neigh_rank.clear(); h->neigh_off.clear();
}

void free_xstage(Handle* h) {
  for (int a = 0; a < 3; a++) {
    cudaFree(h->xs_mask[a]); cudaFree(h->xs_pred[a]);
    h->xs_mask[a] = nullptr; h->xs_pred[a] = nullptr; h->xs_have[a] = false;
  }
  cudaFree(h->xs_off); cudaFree(h->xs_dof); cudaFree(h->xs_skip);
  h->xs_off = h->xs_dof = nullptr; h->xs_skip = nullptr;
  h->xs_valid = false; h->xs_nclass = 0; h->xs_nslots = 0; h->xs_nlinked = 0; h->xs_nmember = 0; h->xs_level = 0;
}

// Staged direct-stiffness summation (gs_kernels.cuh "staged"): per-element face masks / neighbours and the CSR
// lists of the classes left to the class-list pass.  Needs gs_init (and gs_init_shared, if any) done.
int build_xstage(Handle* h, int level) {
  free_xstage(h);
  if (!h->have_gs || h->lx != 8 || h->nelv == 0) return B200_OK;
  cudaStream_t st = h->stream;
  const int nc = h->nclass, ne = h->nelv, threads = 256;
  const int nslots = std::min((ne + XS_NE - 1) / XS_NE, h->num_sm) * XS_NE;
  const int dirs = (level >= 2) ? 7 : 1;
  for (int a = 0; a < 3; a++) {
    if (int r = dmalloc(&h->xs_mask[a], (size_t)ne)) return r;
    if (int r = dmalloc(&h->xs_pred[a], (size_t)ne)) return r;
    CK(cudaMemsetAsync(h->xs_mask[a], 0, sizeof(unsigned long long) * (size_t)ne, st));
    CK(cudaMemsetAsync(h->xs_pred[a], 0xff, sizeof(int) * (size_t)ne, st));
  }
  h->xs_nslots = nslots;
  h->xs_level = level;
  if (nc == 0) {
    if (int r = dmalloc(&h->xs_off, 1)) return r;
    if (int r = dmalloc(&h->xs_dof, 1)) return r;
    CK(cudaMemsetAsync(h->xs_off, 0, sizeof(int), st));
    h->xs_valid = true;
    return B200_OK;
  }
  DevTemps T;
  SgArrays A;
  for (int a = 0; a < 3; a++) {
    if (int r = T.alloc(&A.cnt[a], (size_t)ne)) return r;
    if (int r = T.alloc(&A.pmin[a], (size_t)ne)) return r;
    if (int r = T.alloc(&A.pmax[a], (size_t)ne)) return r;
    if (int r = T.alloc(&A.succ[a], (size_t)ne)) return r;
    if (int r = T.alloc(&A.scnt[a], (size_t)ne)) return r;
    A.pred[a] = h->xs_pred[a];
    A.mask[a] = h->xs_mask[a];
    CK(cudaMemsetAsync(A.cnt[a], 0, sizeof(int) * (size_t)ne, st));
    CK(cudaMemsetAsync(A.pmin[a], 0x7f, sizeof(int) * (size_t)ne, st));      // 0x7f7f7f7f: above every element index
    CK(cudaMemsetAsync(A.pmax[a], 0xff, sizeof(int) * (size_t)ne, st));      // -1
    CK(cudaMemsetAsync(A.succ[a], 0xff, sizeof(int) * (size_t)ne, st));
    CK(cudaMemsetAsync(A.scnt[a], 0, sizeof(int) * (size_t)ne, st));
  }
  int *d_keep = nullptr, *d_mem = nullptr, *d_newidx = nullptr, *d_newoff = nullptr;
  if (int r = T.alloc(&d_keep, (size_t)nc)) return r;
  if (int r = T.alloc(&d_mem, (size_t)nc)) return r;
  if (int r = T.alloc(&d_newidx, (size_t)nc)) return r;
  if (int r = T.alloc(&d_newoff, (size_t)nc)) return r;
  const int gc = grid_for(nc, threads, h->num_sm, 8), ge = grid_for(ne, threads, h->num_sm, 8);
  if (!h->xs_nolink) {
    sg_links_kernel<<<gc, threads, 0, st>>>(h->gs_off, h->gs_dof, h->gs_skip, nc, ne, nslots, dirs, A);
    LAUNCHED();
    sg_elem_kernel<<<ge, threads, 0, st>>>(ne, A);
    LAUNCHED();
    sg_succ_kernel<<<ge, threads, 0, st>>>(ne, A);
    LAUNCHED();
  }
  sg_classify_kernel<<<gc, threads, 0, st>>>(h->gs_off, h->gs_dof, h->gs_skip, nc, dirs, A, d_keep, d_mem);
  LAUNCHED();
  CK(cudaGetLastError());
  void* d_tmp = nullptr;
  size_t tb1 = 0, tb2 = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb1, d_keep, d_newidx, nc, st));
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_mem, d_newoff, nc, st));
  if (int r = T.alloc(reinterpret_cast<unsigned char**>(&d_tmp), std::max(tb1, tb2))) return r;
  CK(cub::DeviceScan::ExclusiveSum(d_tmp, tb1, d_keep, d_newidx, nc, st));
  CK(cub::DeviceScan::ExclusiveSum(d_tmp, tb2, d_mem, d_newoff, nc, st));
  int last[4] = {};
  CK(cudaMemcpyAsync(&last[0], d_newidx + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[1], d_keep + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[2], d_newoff + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[3], d_mem + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  // which directions have any staged node at all (an empty face pass is not launched)
  unsigned long long* d_any = nullptr;
  if (int r = T.alloc(&d_any, 3)) return r;
  for (int a = 0; a < 3; a++) {
    void* d_t2 = nullptr;
    size_t tb = 0;
    CK(cub::DeviceReduce::Max(nullptr, tb, h->xs_mask[a], d_any + a, ne, st));
    if (int r = T.alloc(reinterpret_cast<unsigned char**>(&d_t2), tb)) return r;
    CK(cub::DeviceReduce::Max(d_t2, tb, h->xs_mask[a], d_any + a, ne, st));
  }
  unsigned long long any[3] = {};
  CK(cudaMemcpyAsync(any, d_any, sizeof any, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int a = 0; a < 3; a++) h->xs_have[a] = any[a] != 0ull;
  const int nkeep = last[0] + last[1], nmem = last[2] + last[3];
  if (int r = dmalloc(&h->xs_off, (size_t)nkeep + 1)) return r;
  if (int r = dmalloc(&h->xs_dof, (size_t)nmem)) return r;
  if (h->gs_skip) if (int r = dmalloc(&h->xs_skip, (size_t)nkeep)) return r;
  xs_compact_kernel<<<gc, threads, 0, st>>>(h->gs_off, h->gs_dof, nc, d_keep, d_newidx, d_newoff, h->gs_skip,
                                            h->xs_off, h->xs_dof, h->xs_skip);
  LAUNCHED();
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->xs_off + nkeep, &nmem, sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  h->xs_nclass = nkeep;
  h->xs_nmember = nmem;
  h->xs_nlinked = (int64_t)nc - nkeep;       // classes staged
  h->xs_valid = true;
  return B200_OK;
}

// y and z face passes of the staged summation (after the element kernel, y before z: both touch the i-edges)
int gs_face_passes(Handle* h, double* f0, double* f1, double* f2) {
  if (h->xs_level < 2) return B200_OK;
  const int threads = 256;
  const int grid = grid_for((int64_t)h->nelv * 32, threads, h->num_sm, h->face_ctas);
  if (h->xs_have[1]) {
    gs_face_pass_kernel<1><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->xs_pred[1], h->xs_mask[1], h->nelv);
    LAUNCHED();
  }
  if (h->xs_have[2]) {
    gs_face_pass_kernel<2><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->xs_pred[2], h->xs_mask[2], h->nelv);
    LAUNCHED();
  }
  CK(cudaGetLastError());
  return B200_OK;
}

// separate gather-scatter pass over the packed lists of the schedule (3 fields); entries [lo[b], hi[b]) of
// list b = 0..3 (pairs, quads, octs, hexes)
int gs_packed_range(Handle* h, double* f0, double* f1, double* f2, const int lo[4], const int hi[4]) {
  const int threads = 256;
  if (hi[0] > lo[0]) {
    constexpr int UN = 4;
    const int cnt = hi[0] - lo[0];
    const int grid = grid_for((cnt + UN - 1) / UN, threads, h->num_sm, 8);
    gs_pairs_kernel<UN><<<grid, threads, 0, h->stream>>>(
        f0, f1, f2, reinterpret_cast<const int2*>(h->sched_pair) + lo[0], cnt);
    LAUNCHED();
  }
  if (hi[1] > lo[1]) {
    const int cnt = hi[1] - lo[1];
    gs_wide_kernel<4><<<grid_for(cnt, threads, h->num_sm, 8), threads, 0, h->stream>>>(
        f0, f1, f2, reinterpret_cast<const int4*>(h->sched_quad) + lo[1], cnt);
    LAUNCHED();
  }
  if (hi[2] > lo[2]) {
    const int cnt = hi[2] - lo[2];
    gs_wide_kernel<8><<<grid_for(cnt, threads, h->num_sm, 8), threads, 0, h->stream>>>(
        f0, f1, f2, reinterpret_cast<const int4*>(h->sched_oct) + 2 * (size_t)lo[2], cnt);
    LAUNCHED();
  }
  if (hi[3] > lo[3]) {
    const int cnt = hi[3] - lo[3];
    gs_wide_kernel<16><<<grid_for(cnt, threads, h->num_sm, 4), threads, 0, h->stream>>>(
        f0, f1, f2, reinterpret_cast<const int4*>(h->sched_hex) + 4 * (size_t)lo[3], cnt);
    LAUNCHED();
  }
  CK(cudaGetLastError());
  return B200_OK;
}
int gs_packed(Handle* h, double* f0, double* f1, double* f2) {
  const int lo[4] = {0, 0, 0, 0}, hi[4] = {h->sched_n2, h->sched_n4, h->sched_n8, h->sched_n16};
  return gs_packed_range(h, f0, f1, f2, lo, hi);
}

// classes with more than 16 members are not in the packed lists
int gs_leftover(Handle* h, double* f0, double* f1, double* f2) {
  if (h->sched_nleft == 0) return B200_OK;
  const int threads = 256, grid = grid_for(h->sched_nleft, threads, h->num_sm, 4);
  gs_op_list_kernel<3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->gs_off, h->gs_dof, h->sched_left,
                                                         h->sched_nleft);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

LaunchArgs make_args(const void* vx, const void* vy, const void* vz, const void* vxb, const void* vyb,
                     const void* vzb, const void* rho, const void* chi_in, const void* fsx,
                     const void* fsy, const void* fsz, void* fx, void* fy, void* fz, void* sens,
                     void* chi_out, int nelv) {
  LaunchArgs a;
  a.v[0] = (const double*)vx; a.v[1] = (const double*)vy; a.v[2] = (const double*)vz;
  a.vb[0] = (const double*)vxb; a.vb[1] = (const double*)vyb; a.vb[2] = (const double*)vzb;
  a.rho = rho ? (const double*)rho : (const double*)chi_in;
  a.rho_is_chi = (rho == nullptr);
  a.sources = (a.rho != nullptr);
  a.fs[0] = (const double*)fsx; a.fs[1] = (const double*)fsy; a.fs[2] = (const double*)fsz;
  a.fin[0] = a.fin[1] = a.fin[2] = nullptr;
  a.f[0] = (double*)fx; a.f[1] = (double*)fy; a.f[2] = (double*)fz;
  a.sens = (double*)sens;
  a.chi_out = (double*)chi_out;
  a.elem_list = nullptr;
  a.nelem = nelv;
  a.elem_begin = 0;
  a.no_dealias = false;
  a.xstage = false;
  return a;
}

int masked_lube_post(Handle* h, const LaunchArgs& a) {
  if (!(a.sources && h->if_lube && h->lube_mask_size > 0)) return B200_OK;
  const int threads = 256;
  const int grid = grid_for(h->lube_mask_size, threads, h->num_sm, 4);
  lube_mask_post_kernel<<<grid, threads, 0, h->stream>>>(a.f[0], a.f[1], a.f[2], a.vb[0], a.vb[1], a.vb[2],
                                                         a.rho, h->B, h->K_lube, a.rho_is_chi ? 0 : 1,
                                                         h->convex_up, h->f_min, h->f_max, h->q,
                                                         h->lube_mask, h->lube_mask_size);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int b200_version(void) { return 100; }
void b200_set_abort_on_error(const int* flag) { g_abort = flag ? *flag : 1; }
const char* b200_last_error(void) { return g_err.c_str(); }
int64_t b200_launch_count(void) { return g_launches.load(); }

int b200_adjrhs_create(void** handle, const int* lx, const int* nelv, const int* device) {
  if (!handle || !lx || !nelv) return fail(B200_ERR_ARG, "create: null argument");
  if (*lx < 4 || *lx > LX_MAX) return fail(B200_ERR_ARG, "create: lx=%d outside 4..%d", *lx, LX_MAX);
  if (*nelv < 0) return fail(B200_ERR_ARG, "create: nelv=%d", *nelv);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback (%s)", cudaGetErrorString(e));
  Handle* h = new Handle();
  h->lx = *lx; h->nelv = *nelv; h->n = (int64_t)(*lx) * (*lx) * (*lx) * (*nelv);
  if (h->n > 0x7fffffffll) {
    const long long nn = h->n;
    delete h;
    return fail(B200_ERR_ARG, "create: n=%lld exceeds int32 dof indexing", nn);
  }
  // *device < 0 (or NULL): the calling thread's current CUDA device -- what a Neko rank has selected in
  // device_init; the handle then never moves the thread to another device
  h->device = device ? *device : -1;
  if (h->device < 0) {
    if (cudaGetDevice(&h->device) != cudaSuccess) {
      delete h;
      return fail(B200_ERR_CUDA, "create: cudaGetDevice failed");
    }
  }
  if (h->device >= ndev) {
    const int d = h->device;
    delete h;
    return fail(B200_ERR_ARG, "create: device %d of %d", d, ndev);
  }
  CK(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10)
    fprintf(stderr, "[neko_top_b200] warning: device %s is sm_%d%d, kernels are built for sm_100a\n",
            prop.name, prop.major, prop.minor);
  h->num_sm = prop.multiProcessorCount;
  const char* c = getenv("B200_ADJRHS_CFG");
  h->cfg = c ? atoi(c) : -1;
  const char* g = getenv("B200_GS_MODE");
  if (g) h->gs_mode = atoi(g) != 0 ? 1 : 0;
  g = getenv("B200_EXCHANGE_OVERLAP");
  if (g) h->overlap_elem = (strcmp(g, "elem") == 0);
  g = getenv("B200_EXCHANGE_SIDE");
  if (g) h->exchange_side = atoi(g) != 0;
  g = getenv("B200_PHASE_TIMING");
  if (g) h->phase_timing = atoi(g) != 0;
  g = getenv("B200_GS_UN");
  if (g) h->gs_un = atoi(g);
  g = getenv("B200_XSTAGE");
  if (g) h->xs_enable = std::min(2, std::max(0, atoi(g)));
  g = getenv("B200_FACE_CTAS");
  if (g) h->face_ctas = std::min(8, std::max(1, atoi(g)));
  g = getenv("B200_XS_NOLINK");
  if (g) h->xs_nolink = atoi(g) != 0;
  *handle = h;
  return B200_OK;
}

int b200_adjrhs_free(void** handle) {
  if (!handle || !*handle) return B200_OK;
  Handle* h = H(*handle);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->gs_off); cudaFree(h->gs_dof); cudaFree(h->gs_rep);
  free_shared(h);
  cudaFree(h->geom_pack); cudaFree(h->G_fine);
  free_schedule(h); free_xstage(h); cudaFree(h->d_order); cudaFree(h->work6); cudaFree(h->d_partial);
  cudaFree(h->d_bnd_elem); cudaFree(h->d_int_elem);
  for (double* s : h->stage) cudaFree(s);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->pev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_h2d) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_k) cudaEventDestroy(e);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->ev_pack) cudaEventDestroy(h->ev_pack);
  if (h->ev_recv) cudaEventDestroy(h->ev_recv);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->comm) ncclCommDestroy(h->comm);
  delete h;
  *handle = nullptr;
  return B200_OK;
}

int b200_adjrhs_set_stream(void* handle, void* stream) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  H(handle)->stream = (cudaStream_t)stream;
  return B200_OK;
}

int b200_adjrhs_set_space(void* handle, const double* dx, const double* wx) {
  if (!handle || !dx || !wx) return fail(B200_ERR_ARG, "set_space: null argument");
  Handle* h = H(handle);
  memcpy(h->D, dx, sizeof(double) * h->lx * h->lx);
  memcpy(h->w, wx, sizeof(double) * h->lx);
  h->have_space = true;
  return B200_OK;
}

int b200_adjrhs_set_geometry(void* handle, const void* drdx, const void* dsdx, const void* dtdx,
                             const void* drdy, const void* dsdy, const void* dtdy, const void* drdz,
                             const void* dsdz, const void* dtdz, const void* B) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  const void* g[9] = {drdx, dsdx, dtdx, drdy, dsdy, dtdy, drdz, dsdz, dtdz};
  for (int i = 0; i < 9; i++) {
    if (!g[i]) return fail(B200_ERR_ARG, "set_geometry: null geometric factor %d", i);
    h->G[i] = (const double*)g[i];
  }
  if (!B) return fail(B200_ERR_ARG, "set_geometry: null B");
  h->B = (const double*)B;
  if (int r = check_fields({drdx, dsdx, dtdx, drdy, dsdy, dtdy, drdz, dsdz, dtdz, B})) return r;
  // private per-plane interleaved image for the fused kernel (one bulk copy per plane)
  CK(cudaSetDevice(h->device));
  if (!h->geom_pack && h->n > 0) CK(cudaMalloc(&h->geom_pack, sizeof(double) * NGEO * (size_t)h->n));
  if (h->n > 0) {
    GeomPtrs gp;
    for (int i = 0; i < 9; i++) gp.p[i] = h->G[i];
    gp.p[9] = h->B;
    const int threads = 256;
    geom_pack_kernel<<<grid_for(h->n, threads, h->num_sm, 16), threads, 0, h->stream>>>(
        gp, h->geom_pack, h->lx * h->lx, h->n);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  h->have_geom = true;
  return B200_OK;
}

int b200_adjrhs_set_params(void* handle, const double* f_min, const double* f_max, const double* q,
                           const int* convex_up, const int* if_lube, const double* K_lube,
                           const double* K_sens) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (f_min) h->f_min = *f_min;
  if (f_max) h->f_max = *f_max;
  if (q) h->q = *q;
  if (convex_up) h->convex_up = *convex_up;
  if (if_lube) h->if_lube = *if_lube;
  if (K_lube) h->K_lube = *K_lube;
  if (K_sens) h->K_sens = *K_sens;
  return B200_OK;
}

int b200_adjrhs_set_lube_mask(void* handle, const void* mask_d, const int* mask_size) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  h->lube_mask = (const int*)mask_d;
  h->lube_mask_size = (mask_d && mask_size) ? *mask_size : 0;
  return B200_OK;
}

int b200_adjrhs_compute(void* handle, const void* vx, const void* vy, const void* vz, const void* vxb,
                        const void* vyb, const void* vzb, const void* rho, const void* chi_in,
                        const void* fsx, const void* fsy, const void* fsz, void* fx, void* fy, void* fz,
                        void* sens, void* chi_out) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "compute: null velocity / rhs pointer");
  if ((fsx || fsy || fsz) && !(fsx && fsy && fsz)) return fail(B200_ERR_ARG, "compute: partial f_static");
  if (int r = check_fields({vx, vy, vz, vxb, vyb, vzb, rho, chi_in, fsx, fsy, fsz, fx, fy, fz, sens, chi_out}))
    return r;
  CK(cudaSetDevice(h->device));
  NvtxRange nvtx("Fluid: adjoint RHS (b200_adjrhs_compute)");
  LaunchArgs a = make_args(vx, vy, vz, vxb, vyb, vzb, rho, chi_in, fsx, fsy, fsz, fx, fy, fz, sens,
                           chi_out, h->nelv);
  if (int r = launch_fused(h, a)) return r;
  return masked_lube_post(h, a);
}

int b200_adjrhs_step(void* handle, const void* vx, const void* vy, const void* vz, const void* vxb,
                     const void* vyb, const void* vzb, const void* rho, const void* chi_in,
                     const void* fsx, const void* fsy, const void* fsz, void* fx, void* fy, void* fz,
                     void* sens, void* chi_out) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!h->have_gs) return fail(B200_ERR_STATE, "step: b200_gs_init not called");
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "step: null velocity / rhs pointer");
  if (int r = check_fields({vx, vy, vz, vxb, vyb, vzb, rho, chi_in, fsx, fsy, fsz, fx, fy, fz, sens, chi_out}))
    return r;
  CK(cudaSetDevice(h->device));
  LaunchArgs a = make_args(vx, vy, vz, vxb, vyb, vzb, rho, chi_in, fsx, fsy, fsz, fx, fy, fz, sens,
                           chi_out, h->nelv);
  double *f0 = a.f[0], *f1 = a.f[1], *f2 = a.f[2];
  NvtxRange nvtx_step("Fluid: adjoint RHS + Velocity residual gs_op (b200_adjrhs_step)");
  if (int r = time_mark(h)) return r;
  // Multi-GPU, default: ONE launch over all elements, then the classes holding shared nodes are summed and
  // packed, and the NCCL exchange runs on the communication stream WHILE the remaining (rank-local) classes
  // are summed.  The alternative (B200_EXCHANGE_OVERLAP=elem; needs b200_adjrhs_set_boundary_elements)
  // computes the partition-boundary elements first and overlaps the exchange with the interior-element
  // kernel; measured on 4 and 8 B200 it loses: the persistent element kernel owns every SM, so it either
  // delays the NCCL kernel or starts some of its CTAs late behind it (static element partition), and the
  // element-list variant of the kernel is ~12 % slower than the contiguous one.
  const bool mgpu = h->comm && h->nshared > 0 && h->n_shared_cls >= 0 && h->gs_skip;
  // x stage (lx = 8 default kernel, mesh order, CSR pass): the element kernel sums the i-face pair classes of
  // consecutive elements itself and the pass runs over the remaining classes only.  The masked lube term is
  // added to single nodes after the element kernel, so it needs the un-summed values.
  const bool masked_lube = a.sources && h->if_lube && h->lube_mask_size > 0;
  bool xs = h->xs_enable && uses_v3(h) && h->gs_mode == 0 && !h->d_order && !masked_lube &&
            !(mgpu && h->overlap_elem) && !(h->comm && h->nshared > 0 && !mgpu);
  if (xs && (!h->xs_valid || h->xs_level != h->xs_enable)) {
    if (int r = build_xstage(h, h->xs_enable)) return r;
    xs = h->xs_valid;
  }
  if (mgpu && !h->overlap_elem) {
    if (int r = phase_mark(h, 0)) return r;
    a.elem_list = h->d_order;
    a.xstage = xs;
    if (int r = launch_fused(h, a)) return r;
    if (int r = masked_lube_post(h, a)) return r;
    if (int r = phase_mark(h, 1)) return r;
    if (int r = time_mark(h)) return r;
    // shared-class sum and pack, then the NCCL exchange on the communication stream while the main stream runs the
    // rank-local passes (B200_EXCHANGE_SIDE=1 moves sum and pack to the communication stream too: measured
    // 4.657 vs 4.629 ms at N = 2, r02n -- no gain)
    if (int r = phase_mark(h, 2)) return r;
    if (int r = gs_exchange(h, f0, f1, f2, 3, true, h->exchange_side)) return r;
    if (int r = phase_mark(h, 3)) return r;
    if (xs) if (int r = gs_face_passes(h, f0, f1, f2)) return r;
    if (int r = gs_step_pass(h, f0, f1, f2, xs)) return r;
    if (int r = phase_mark(h, 4)) return r;
    if (int r = gs_finish_exchange(h, f0, f1, f2, 3)) return r;
    if (int r = phase_mark(h, 5)) return r;
    return time_mark(h);
  }
  const bool split = mgpu && h->nbnd > 0;
  // direct-stiffness summation inside the element kernel (v3, lx = 8) while f is still in L2; the masked
  // lube term is applied by a separate kernel after the element kernel, so it keeps the separate gs pass
  // gs_mode 1: separate pass over the packed class lists of the
  // schedule; 0: the CSR kernels.  Without a boundary/interior split a communicator needs the CSR pass
  // (it sums the shared classes too, before the exchange).
  int mode = h->gs_mode;
  if (h->nclass == 0 || (!split && h->comm && h->nshared > 0)) mode = 0;
  if (mode > 0) {
    const int kind = split ? 2 : 1;
    if (!h->sched_valid || h->sched_kind != kind) {
      if (int r = build_gs_schedule(h, split ? h->d_int_elem : h->d_order, split ? h->nint : h->nelv)) return r;
      h->sched_kind = kind;
    }
    if (!h->sched_valid) mode = 0;
  }
  if (split) {
    // boundary elements first, their shared nodes summed locally, packed and sent while the interior
    // elements are computed
    // (with a masked lube term the point-zone kernel must see every element before any summation: no overlap)
    LaunchArgs ab = a;
    if (!masked_lube) { ab.elem_list = h->d_bnd_elem; ab.nelem = h->nbnd; }
    if (int r = phase_mark(h, 0)) return r;
    if (int r = launch_fused(h, ab)) return r;
    if (int r = masked_lube_post(h, a)) return r;
    if (int r = phase_mark(h, 1)) return r;
    if (h->n_shared_cls > 0) {
      const int threads = 256, grid = grid_for(h->n_shared_cls, threads, h->num_sm, 4);
      gs_op_list_kernel<3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->gs_off, h->gs_dof,
                                                             h->gs_shared_cls, h->n_shared_cls);
      LAUNCHED();
      CK(cudaGetLastError());
    }
    if (int r = phase_mark(h, 2)) return r;
    if (int r = gs_exchange(h, f0, f1, f2, 3)) return r;
    if (int r = phase_mark(h, 3)) return r;
    LaunchArgs ai = a; ai.elem_list = h->d_int_elem; ai.nelem = h->nint;
    if (h->nint > 0 && !masked_lube) if (int r = launch_fused(h, ai)) return r;
    if (int r = phase_mark(h, 4)) return r;
    if (int r = time_mark(h)) return r;
    if (mode == 1) {
      if (int r = gs_packed(h, f0, f1, f2)) return r;
      if (int r = gs_leftover(h, f0, f1, f2)) return r;
    } else if (h->nclass > 0) {
      const int threads = 256, grid = grid_for(h->nclass, threads, h->num_sm, 3);
      gs_op_kernel<3, 1, 3><<<grid, threads, 0, h->stream>>>(f0, f1, f2, h->gs_off, h->gs_dof, h->nclass,
                                                              h->gs_skip);
      LAUNCHED();
      CK(cudaGetLastError());
    }
    if (int r = phase_mark(h, 5)) return r;
    if (int r = gs_finish_exchange(h, f0, f1, f2, 3)) return r;
    if (int r = phase_mark(h, 6)) return r;
  } else {
    a.elem_list = h->d_order;
    a.xstage = xs && mode == 0;
    if (int r = launch_fused(h, a)) return r;
    if (int r = masked_lube_post(h, a)) return r;
    if (int r = time_mark(h)) return r;
    if (mode == 1) {
      if (int r = gs_packed(h, f0, f1, f2)) return r;
      if (int r = gs_leftover(h, f0, f1, f2)) return r;
    } else if (a.xstage) {
      if (int r = gs_face_passes(h, f0, f1, f2)) return r;
      if (int r = gs_step_pass(h, f0, f1, f2, true)) return r;
    } else {
      if (int r = gs_launch(h, f0, f1, f2, 3)) return r;
      if (int r = gs_exchange(h, f0, f1, f2, 3)) return r;
      if (int r = gs_finish_exchange(h, f0, f1, f2, 3)) return r;
    }
  }
  return time_mark(h);
}

int b200_adjrhs_set_element_order(void* handle, const int* nelem, const int* order) {
  if (!handle || !nelem) return fail(B200_ERR_ARG, "set_element_order: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  free_schedule(h);
  cudaFree(h->d_order); h->d_order = nullptr;
  h->order.clear();
  if (*nelem == 0 || !order) return B200_OK;     // back to 0..nelv-1
  if (*nelem != h->nelv) return fail(B200_ERR_ARG, "set_element_order: %d entries, nelv = %d", *nelem, h->nelv);
  std::vector<char> seen(h->nelv, 0);
  for (int i = 0; i < h->nelv; i++) {
    const int e = order[i];
    if (e < 0 || e >= h->nelv || seen[e]) return fail(B200_ERR_ARG, "set_element_order: not a permutation");
    seen[e] = 1;
  }
  h->order.assign(order, order + h->nelv);
  if (int r = dmalloc(&h->d_order, (size_t)h->nelv)) return r;
  CK(cudaMemcpy(h->d_order, order, sizeof(int) * (size_t)h->nelv, cudaMemcpyHostToDevice));
  if (h->nbnd > 0 && h->d_bnd_elem) {            // keep the interior list in the new order
    std::vector<int> bnd(h->nbnd);
    CK(cudaMemcpy(bnd.data(), h->d_bnd_elem, sizeof(int) * (size_t)h->nbnd, cudaMemcpyDeviceToHost));
    const int nb = h->nbnd;
    return b200_adjrhs_set_boundary_elements(handle, &nb, bnd.data());
  }
  return B200_OK;
}

int b200_adjrhs_set_gs_fused(void* handle, const int* flag) {
  if (!handle || !flag) return fail(B200_ERR_ARG, "set_gs_fused: null argument");
  H(handle)->gs_mode = (*flag != 0) ? 1 : 0;
  return B200_OK;
}

int b200_adjrhs_gs_info(void* handle, int* fused, int64_t* classes_in_kernel, int64_t* classes_total) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (fused) *fused = 0;       // the in-kernel class summation of round 1 is gone (superseded by the staged summation)
  if (classes_in_kernel) *classes_in_kernel = (h->sched_valid && h->gs_mode == 1) ? h->sched_nfused : 0;
  if (classes_total) *classes_total = h->nclass;
  return B200_OK;
}

int b200_adjrhs_set_xstage(void* handle, const int* flag) {
  if (!handle || !flag) return fail(B200_ERR_ARG, "set_xstage: null argument");
  H(handle)->xs_enable = std::min(2, std::max(0, *flag));
  return B200_OK;
}

int b200_adjrhs_xstage_info(void* handle, int* active, int64_t* classes_staged, int64_t* classes_left,
                            int64_t* classes_total) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  const bool on = h->xs_enable && h->xs_valid && h->xs_level == h->xs_enable;
  if (active) *active = on ? h->xs_level : 0;
  if (classes_staged) *classes_staged = on ? h->xs_nlinked : 0;
  if (classes_left) *classes_left = on ? h->xs_nclass : h->nclass;
  if (classes_total) *classes_total = h->nclass;
  return B200_OK;
}

int b200_adv_adjoint_compute(void* handle, const void* vx, const void* vy, const void* vz,
                             const void* vxb, const void* vyb, const void* vzb, void* fx, void* fy,
                             void* fz) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "adv_adjoint_compute: null pointer");
  if (int r = check_fields({vx, vy, vz, vxb, vyb, vzb, fx, fy, fz})) return r;
  CK(cudaSetDevice(h->device));
  LaunchArgs a = make_args(vx, vy, vz, vxb, vyb, vzb, nullptr, nullptr, nullptr, nullptr, nullptr, fx, fy,
                           fz, nullptr, nullptr, h->nelv);
  a.fin[0] = (const double*)fx; a.fin[1] = (const double*)fy; a.fin[2] = (const double*)fz;
  a.no_dealias = true;
  return launch_fused(h, a);
}

int b200_adv_linear_compute(void* handle, const void* vx, const void* vy, const void* vz,
                            const void* vxb, const void* vyb, const void* vzb, const void* jacinv,
                            void* fx, void* fy, void* fz) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "adv_linear_compute: null pointer");
  (void)jacinv;   // B*jacinv == w3 (coef%B = jac*w3): the kernel uses the quadrature weights directly
  CK(cudaSetDevice(h->device));
  LaunchArgs a = make_args(vx, vy, vz, vxb, vyb, vzb, nullptr, nullptr, nullptr, nullptr, nullptr, fx, fy,
                           fz, nullptr, nullptr, h->nelv);
  a.fin[0] = (const double*)fx; a.fin[1] = (const double*)fy; a.fin[2] = (const double*)fz;
  return launch_fine(h, a, ADV_LINEAR, false);
}

int b200_adv_dealias_init(void* handle, const int* lxd, const double* interp, const double* dxd,
                          const double* wd) {
  if (!handle || !lxd || !interp || !dxd || !wd) return fail(B200_ERR_ARG, "adv_dealias_init: null argument");
  Handle* h = H(handle);
  if (!h->have_geom) return fail(B200_ERR_STATE, "adv_dealias_init: call b200_adjrhs_set_geometry first");
  const int ld = *lxd;
  if (ld != advop_default_lxd(h->lx))
    return fail(B200_ERR_ARG, "adv_dealias_init: lxd=%d, only 3*lx/2=%d is instantiated for lx=%d "
                "(advection_adjoint_fctry.f90:70,89)", ld, advop_default_lxd(h->lx), h->lx);
  CK(cudaSetDevice(h->device));
  h->Jd.assign(interp, interp + (size_t)ld * h->lx);
  h->Dd.assign(dxd, dxd + (size_t)ld * ld);
  h->wd.assign(wd, wd + ld);
  const size_t nd = (size_t)h->nelv * ld * ld * ld;
  if (h->G_fine) { CK(cudaFree(h->G_fine)); h->G_fine = nullptr; }
  CK(cudaMalloc(&h->G_fine, sizeof(double) * 9 * std::max<size_t>(nd, 1)));
  double* dst[9];
  for (int g = 0; g < 9; g++) dst[g] = h->G_fine + g * nd;
  const char* msg = nullptr;
  cudaError_t e = advop_geom_to_fine(h->lx, ld, h->Jd.data(), h->G, dst, h->nelv, h->num_sm, h->stream, &msg);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "adv_dealias_init: %s", msg ? msg : cudaGetErrorString(e));
  if (h->nelv > 0) LAUNCHED();
  CK(cudaStreamSynchronize(h->stream));
  h->lxd = ld;
  return B200_OK;
}

int b200_adjrhs_set_dealias(void* handle, const int* flag) {
  if (!handle || !flag) return fail(B200_ERR_ARG, "set_dealias: null argument");
  Handle* h = H(handle);
  if (*flag && h->lxd == 0) return fail(B200_ERR_STATE, "set_dealias: b200_adv_dealias_init not called");
  h->dealias_fused = (*flag != 0);
  return B200_OK;
}

static int adv_dealias_compute(void* handle, int mode, const void* vx, const void* vy, const void* vz,
                               const void* vxb, const void* vyb, const void* vzb, void* fx, void* fy,
                               void* fz) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "adv dealias compute: null pointer");
  CK(cudaSetDevice(h->device));
  LaunchArgs a = make_args(vx, vy, vz, vxb, vyb, vzb, nullptr, nullptr, nullptr, nullptr, nullptr, fx, fy,
                           fz, nullptr, nullptr, h->nelv);
  a.fin[0] = (const double*)fx; a.fin[1] = (const double*)fy; a.fin[2] = (const double*)fz;
  return launch_fine(h, a, mode, true);
}

int b200_adv_adjoint_dealias_compute(void* handle, const void* vx, const void* vy, const void* vz,
                                     const void* vxb, const void* vyb, const void* vzb, void* fx, void* fy,
                                     void* fz) {
  return adv_dealias_compute(handle, ADV_ADJOINT, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz);
}

int b200_adv_linear_dealias_compute(void* handle, const void* vx, const void* vy, const void* vz,
                                    const void* vxb, const void* vyb, const void* vzb, void* fx, void* fy,
                                    void* fz) {
  return adv_dealias_compute(handle, ADV_LINEAR, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz);
}

// ---- un-fused point-wise drop-ins -------------------------------------------------------------
static int dev_sm_count() {
  static int sm = 0;
  if (!sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    if (sm <= 0) sm = 148;
  }
  return sm;
}

int b200_brinkman_compute(void* fu, void* fv, void* fw, const void* u, const void* v, const void* w,
                          const void* chi, const int* n, void* stream) {
  if (!fu || !fv || !fw || !u || !v || !w || !chi || !n) return fail(B200_ERR_ARG, "brinkman: null argument");
  const int threads = 256;
  if (aligned16(fu) && aligned16(fv) && aligned16(fw) && aligned16(u) && aligned16(v) && aligned16(w) &&
      aligned16(chi)) {
    brinkman_kernel<<<grid_for((*n + 1) / 2, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
        (double*)fu, (double*)fv, (double*)fw, (const double*)u, (const double*)v, (const double*)w,
        (const double*)chi, *n);
  } else {
    // -1*chi*u == lube kernel with K = -1 (scalar path for unaligned views)
    lube_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
        (double*)fu, (double*)fv, (double*)fw, (const double*)u, (const double*)v, (const double*)w,
        (const double*)chi, -1.0, *n);
  }
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_lube_compute(void* fu, void* fv, void* fw, const void* u, const void* v, const void* w,
                      const void* chi, const double* K, const void* mask_d, const int* mask_size,
                      const int* n, void* stream) {
  if (!fu || !fv || !fw || !u || !v || !w || !chi || !K || !n) return fail(B200_ERR_ARG, "lube: null argument");
  const int threads = 256;
  const int ms = (mask_d && mask_size) ? *mask_size : 0;
  if (ms > 0)
    lube_mask_kernel<<<grid_for(ms, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
        (double*)fu, (double*)fv, (double*)fw, (const double*)u, (const double*)v, (const double*)w,
        (const double*)chi, *K, (const int*)mask_d, ms);
  else
    lube_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
        (double*)fu, (double*)fv, (double*)fw, (const double*)u, (const double*)v, (const double*)w,
        (const double*)chi, *K, *n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_opcolv(void* fx, void* fy, void* fz, const void* B, const int* n, void* stream) {
  if (!fx || !fy || !fz || !B || !n) return fail(B200_ERR_ARG, "opcolv: null argument");
  const int threads = 256;
  opcolv_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      (double*)fx, (double*)fy, (double*)fz, (const double*)B, *n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_ramp_forward(void* chi, const void* rho, const int* n, const double* f_min, const double* f_max,
                      const double* q, const int* convex_up, void* stream) {
  if (!chi || !rho || !n || !f_min || !f_max || !q || !convex_up) return fail(B200_ERR_ARG, "ramp: null argument");
  const int threads = 256;
  ramp_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      (double*)chi, (const double*)rho, *n, *f_min, *f_max, *q, *convex_up);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_ramp_backward(void* dF_drho, const void* dF_dchi, const void* rho, const int* n,
                       const double* f_min, const double* f_max, const double* q, const int* convex_up,
                       void* stream) {
  if (!dF_drho || !dF_dchi || !rho || !n || !f_min || !f_max || !q || !convex_up)
    return fail(B200_ERR_ARG, "ramp_backward: null argument");
  const int threads = 256;
  ramp_backward_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      (double*)dF_drho, (const double*)dF_dchi, (const double*)rho, *n, *f_min, *f_max, *q, *convex_up);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_sensitivity(void* sens, const void* u, const void* v, const void* w, const void* ua,
                     const void* va, const void* wa, const double* K_obj, const int* if_lube,
                     const int* n, void* stream) {
  if (!sens || !u || !v || !w || !ua || !va || !wa || !K_obj || !if_lube || !n)
    return fail(B200_ERR_ARG, "sensitivity: null argument");
  const int threads = 256;
  sensitivity_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      (double*)sens, (const double*)u, (const double*)v, (const double*)w, (const double*)ua,
      (const double*)va, (const double*)wa, *K_obj, *if_lube, *n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_steady_field_update(double* result, const void* x, void* x_old, const int* n, void* stream) {
  if (!result || !n) return fail(B200_ERR_ARG, "steady_field_update: null argument");
  if (*n < 0) return fail(B200_ERR_ARG, "steady_field_update: n=%d", *n);
  if (*n == 0) { *result = 0.0; return B200_OK; }      // a rank without elements
  if (!x || !x_old) return fail(B200_ERR_ARG, "steady_field_update: null field");
  // per-device scratch (partials + result), allocated once: the call runs every time step
  static double* scratch[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(B200_ERR_ARG, "steady_field_update: device index %d", dev);
  if (!scratch[dev]) CK(cudaMalloc(&scratch[dev], sizeof(double) * (STEADY_BLOCKS + 1)));
  double* d_part = scratch[dev];
  double* d_res = d_part + STEADY_BLOCKS;
  cudaStream_t st = (cudaStream_t)stream;
  steady_update_kernel<<<STEADY_BLOCKS, 256, 0, st>>>(d_part, (const double*)x, (double*)x_old, (int64_t)*n);
  LAUNCHED();
  steady_reduce_kernel<<<1, 256, 0, st>>>(d_part, d_res);
  LAUNCHED();
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(result, d_res, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return B200_OK;
}

// ---- minimum-dissipation objective chain -------------------------------------------------
static int deriv_run(Handle* h, int mode, const void* u1, const void* u2, const void* u3, const void* jacinv,
                     double* o0, double* o1, double* o2) {
  DerivLaunch L;
  memset(&L, 0, sizeof L);
  L.lx = h->lx; L.mode = mode; L.nelv = h->nelv; L.D = h->D;
  L.u[0] = (const double*)u1; L.u[1] = (const double*)u2; L.u[2] = (const double*)u3;
  for (int g = 0; g < 9; g++) L.G[g] = h->G[g];
  L.jacinv = (const double*)jacinv; L.B = h->B;
  L.out[0] = o0; L.out[1] = o1; L.out[2] = o2;
  L.num_sm = h->num_sm; L.stream = h->stream;
  const char* msg = nullptr;
  cudaError_t e = deriv_launch(L, &msg);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "derivative kernel: %s", msg ? msg : cudaGetErrorString(e));
  if (h->nelv > 0) LAUNCHED();
  return B200_OK;
}

// deterministic local dot product on the handle's stream (synchronises: returns a host scalar)
static int dot_local(Handle* h, const double* a, const double* b, const int* mask, int mask_size, double* result) {
  const int nblk = 1024;
  if (!h->d_partial) if (int r = dmalloc(&h->d_partial, (size_t)nblk + 64)) return r;
  const int64_t count = mask ? (int64_t)mask_size : h->n;
  dot_partial_kernel<<<nblk, 256, 0, h->stream>>>(a, b, mask, count, h->d_partial);
  LAUNCHED();
  CK(cudaGetLastError());
  std::vector<double> part(nblk);
  CK(cudaMemcpyAsync(part.data(), h->d_partial, sizeof(double) * nblk, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double s = 0.0;
  for (double v : part) s += v;
  *result = s;
  return B200_OK;
}

int b200_curl(void* handle, void* w1, void* w2, void* w3, const void* u1, const void* u2, const void* u3,
              const void* jacinv, const void* Binv) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!w1 || !w2 || !w3 || !u1 || !u2 || !u3 || !jacinv || !Binv) return fail(B200_ERR_ARG, "curl: null argument");
  if (!h->have_space || !h->have_geom) return fail(B200_ERR_STATE, "set_space/set_geometry not called");
  if (!h->have_gs) return fail(B200_ERR_STATE, "curl: b200_gs_init not called (Neko's curl averages with gs)");
  CK(cudaSetDevice(h->device));
  if (int r = deriv_run(h, DERIV_CURL_B, u1, u2, u3, jacinv, (double*)w1, (double*)w2, (double*)w3)) return r;
  if (int r = b200_gs_op3(handle, w1, w2, w3)) return r;
  const int threads = 256;
  opcolv_kernel<<<grid_for(h->n, threads, h->num_sm, 8), threads, 0, h->stream>>>(
      (double*)w1, (double*)w2, (double*)w3, (const double*)Binv, h->n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_curlcurl_forcing(void* handle, void* fu, void* fv, void* fw, const void* u, const void* v,
                          const void* w, const void* jacinv, const void* Binv, const void* mask_d,
                          const int* mask_size, const double* obj_scale) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!fu || !fv || !fw || !u || !v || !w || !jacinv || !Binv || !obj_scale)
    return fail(B200_ERR_ARG, "curlcurl_forcing: null argument");
  CK(cudaSetDevice(h->device));
  if (!h->work6 && h->n > 0) if (int r = dmalloc(&h->work6, 6 * (size_t)h->n)) return r;
  double* wo[6];
  for (int i = 0; i < 6; i++) wo[i] = h->work6 + (size_t)i * h->n;
  if (int r = b200_curl(handle, wo[0], wo[1], wo[2], u, v, w, jacinv, Binv)) return r;
  if (int r = b200_curl(handle, wo[3], wo[4], wo[5], wo[0], wo[1], wo[2], jacinv, Binv)) return r;
  const int ms = (mask_d && mask_size) ? *mask_size : 0;
  const int64_t count = ms > 0 ? ms : h->n;
  const int threads = 256;
  add2s2_mask3_kernel<<<grid_for(count, threads, h->num_sm, 8), threads, 0, h->stream>>>(
      (double*)fu, (double*)fv, (double*)fw, wo[3], wo[4], wo[5], *obj_scale, ms > 0 ? (const int*)mask_d : nullptr,
      count);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_min_dissipation_objective(void* handle, const void* u, const void* v, const void* w, const void* chi,
                                   const void* jacinv, const void* mask_d, const int* mask_size,
                                   const double* K, const double* obj_scale, double* out3) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!u || !v || !w || !jacinv || !K || !obj_scale || !out3)
    return fail(B200_ERR_ARG, "min_dissipation_objective: null argument");
  if (!h->have_space || !h->have_geom) return fail(B200_ERR_STATE, "set_space/set_geometry not called");
  CK(cudaSetDevice(h->device));
  if (!h->work6 && h->n > 0) if (int r = dmalloc(&h->work6, 6 * (size_t)h->n)) return r;
  double* obj = h->work6;
  const int ms = (mask_d && mask_size) ? *mask_size : 0;
  const int* mask = ms > 0 ? (const int*)mask_d : nullptr;
  if (int r = deriv_run(h, DERIV_DISSIPATION, u, v, w, jacinv, obj, nullptr, nullptr)) return r;
  double diss = 0.0, lube = 0.0;
  if (int r = dot_local(h, obj, h->B, mask, ms, &diss)) return r;
  if (chi) {
    const int threads = 256;
    lube_density_kernel<<<grid_for(h->n, threads, h->num_sm, 8), threads, 0, h->stream>>>(
        obj, (const double*)u, (const double*)v, (const double*)w, (const double*)chi, h->n);
    LAUNCHED();
    CK(cudaGetLastError());
    if (int r = dot_local(h, obj, h->B, mask, ms, &lube)) return r;
  }
  out3[1] = diss; out3[2] = lube;
  out3[0] = (diss + (chi ? 0.5 * (*K) * lube : 0.0)) * (*obj_scale);
  return B200_OK;
}

int b200_mask_exterior_const(void* fld, void* work, const void* mask_d, const int* mask_size, const double* c,
                             const int* n, void* stream) {
  if (!fld || !work || !mask_d || !mask_size || !c || !n) return fail(B200_ERR_ARG, "mask_exterior_const: null argument");
  const int threads = 256;
  cudaStream_t st = (cudaStream_t)stream;
  fill_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, st>>>((double*)work, *c, (int64_t)*n);
  LAUNCHED();
  if (*mask_size > 0) {
    copy_mask_kernel<<<grid_for(*mask_size, threads, dev_sm_count(), 8), threads, 0, st>>>(
        (double*)work, (const double*)fld, (const int*)mask_d, *mask_size);
    LAUNCHED();
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(fld, work, sizeof(double) * (size_t)*n, cudaMemcpyDeviceToDevice, st));
  return B200_OK;
}

// ---- PDE (Helmholtz) filter: SURVEY.md 8f row 4 ------------------------------------------------------------
static int helm_run(Handle* h, int mode, const double* u, const double* jacinv, double* out, double h1, double h2) {
  HelmLaunch L;
  memset(&L, 0, sizeof L);
  L.lx = h->lx; L.mode = mode; L.nelv = h->nelv; L.D = h->D; L.w = h->w; L.u = u;
  for (int g = 0; g < 9; g++) L.G[g] = h->G[g];
  L.jacinv = jacinv; L.B = h->B; L.out = out; L.h1 = h1; L.h2 = h2;
  L.num_sm = h->num_sm; L.stream = h->stream;
  const char* msg = nullptr;
  cudaError_t e = helm_launch(L, &msg);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "Helmholtz kernel: %s", msg ? msg : cudaGetErrorString(e));
  if (h->nelv > 0) LAUNCHED();
  return B200_OK;
}

// glsc3(a, mult, b) into the device state: deterministic block partials, fixed-tree sum, sum over the ranks
// (Neko: MPI_Allreduce; here ncclAllReduce on the device scalar), then the scalar recurrence of `step`
static int cg_dot_step(Handle* h, const double* a, const double* m, const double* b, CgState* st, int step, int it,
                       double nf, double tol) {
  const int nblk = 1024;
  dot3_partial_kernel<<<nblk, 256, 0, h->stream>>>(a, m, b, h->n, h->d_partial);
  LAUNCHED();
  cg_sum_kernel<<<1, 256, 0, h->stream>>>(h->d_partial, nblk, st);
  LAUNCHED();
  if (h->comm && h->nranks > 1) NK(ncclAllReduce(&st->tmp, &st->tmp, 1, ncclDouble, ncclSum, h->comm, h->stream));
  cg_step_kernel<<<1, 32, 0, h->stream>>>(st, step, it, nf, tol);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_pde_filter_apply(void* handle, void* x_out, const void* x_in, const void* jacinv, const void* mult,
                          const double* radius, const double* abs_tol, const int* max_iter,
                          const int* precond, const double* norm_fac, const int* x0_is_input, int* iters,
                          double* res_start, double* res_final) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!x_out || !x_in || !jacinv || !mult || !radius || !abs_tol || !max_iter)
    return fail(B200_ERR_ARG, "pde_filter_apply: null argument");
  if (!h->have_space || !h->have_geom) return fail(B200_ERR_STATE, "set_space/set_geometry not called");
  if (!h->have_gs) return fail(B200_ERR_STATE, "pde_filter_apply: b200_gs_init not called");
  CK(cudaSetDevice(h->device));
  NvtxRange nvtx("filter solve (b200_pde_filter_apply)");     // the reference's region name, PDE_filter_mapping.f90:255
  const int64_t n = h->n;
  if (!h->work6 && n > 0) if (int r = dmalloc(&h->work6, 6 * (size_t)n)) return r;
  if (!h->d_partial) if (int r = dmalloc(&h->d_partial, (size_t)1024 + 64)) return r;
  CgState* st = reinterpret_cast<CgState*>(h->d_partial + 1024);
  double *rr = h->work6, *pp = rr + n, *zz = pp + n, *ww = zz + n, *dinv = ww + n;
  double* x = (double*)x_out;
  const double* mlt = (const double*)mult;
  const double h1 = (*radius) * (*radius), h2 = 1.0;          // PDE_filter_mapping.f90:229-231 / :240-244
  const double nf = norm_fac ? *norm_fac : 1.0;
  const bool jacobi = !precond || *precond != 0;
  const bool x0 = x0_is_input && *x0_is_input != 0;
  const int threads = 256;
  const int grid = grid_for(n, threads, h->num_sm, 8);
  cudaStream_t sm = h->stream;
  // RHS = B * X_in, direct-stiffness summed (:231,251)
  cg_col3_kernel<<<grid, threads, 0, sm>>>(rr, (const double*)x_in, h->B, n);
  LAUNCHED();
  if (int r = b200_gs_op(handle, rr)) return r;
  CK(cudaMemsetAsync(pp, 0, sizeof(double) * (size_t)n, sm));
  if (x0) {
    // "copy the unfiltered design as an initial guess" (:246-248): x = X_in, r = b - A x
    if (x != (const double*)x_in) CK(cudaMemcpyAsync(x, x_in, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, sm));
    if (int r = helm_run(h, 0, x, (const double*)jacinv, ww, h1, h2)) return r;
    if (int r = b200_gs_op(handle, ww)) return r;
    cg_sub_kernel<<<grid, threads, 0, sm>>>(rr, ww, n);
    LAUNCHED();
  } else {
    // Neko's Krylov solvers zero x on entry (the field_copy at :248 is then overwritten)
    CK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, sm));
  }
  if (jacobi) {
    if (int r = helm_run(h, 1, nullptr, (const double*)jacinv, dinv, h1, h2)) return r;
    if (int r = b200_gs_op(handle, dinv)) return r;
    cg_invert_kernel<<<grid, threads, 0, sm>>>(dinv, n);
    LAUNCHED();
  }
  if (int r = cg_dot_step(h, rr, mlt, rr, st, CG_INIT, 0, nf, *abs_tol)) return r;
  CgState hs;
  memset(&hs, 0, sizeof hs);
  const int CHUNK = 8;                    // iterations enqueued between two looks at the state
  int it = 0;
  while (true) {
    CK(cudaMemcpyAsync(&hs, st, sizeof hs, cudaMemcpyDeviceToHost, sm));
    CK(cudaStreamSynchronize(sm));
    if (hs.done || it >= *max_iter) break;
    const int upto = std::min(*max_iter, it + CHUNK);
    for (it = it + 1; it <= upto; it++) {
      const double* z = rr;
      if (jacobi) {
        cg_col3_kernel<<<grid, threads, 0, sm>>>(zz, rr, dinv, n);
        LAUNCHED();
        z = zz;
      }
      if (int r = cg_dot_step(h, rr, mlt, z, st, CG_RTZ, it, nf, *abs_tol)) return r;
      cg_p_update_kernel<<<grid, threads, 0, sm>>>(pp, z, st, n);
      LAUNCHED();
      if (int r = helm_run(h, 0, pp, (const double*)jacinv, ww, h1, h2)) return r;
      if (int r = b200_gs_op(handle, ww)) return r;
      if (int r = cg_dot_step(h, ww, mlt, pp, st, CG_PAP, it, nf, *abs_tol)) return r;
      cg_xr_update_kernel<<<grid, threads, 0, sm>>>(x, rr, pp, ww, st, n);
      LAUNCHED();
      if (int r = cg_dot_step(h, rr, mlt, rr, st, CG_RTR, it, nf, *abs_tol)) return r;
    }
    it = upto;
  }
  CK(cudaGetLastError());
  if (iters) *iters = hs.iters;
  if (res_start) *res_start = hs.res_start;
  if (res_final) *res_final = hs.rnorm;
  return B200_OK;
}

// ---- explicit time scheme around the RHS ---------------------------------------------------------------
static int abf_bdf_launch(void* fx, void* fy, void* fz, void* abx1, void* aby1, void* abz1, void* abx2,
                          void* aby2, void* abz2, int do_abf, const double* ext, int do_bdf, const void* u,
                          const void* v, const void* w, const void* ul1, const void* vl1, const void* wl1,
                          const void* ul2, const void* vl2, const void* wl2, const void* B, double rho,
                          double dt, const double* bd, int nbd, int n, void* stream) {
  Vec3Ptr f{{(double*)fx, (double*)fy, (double*)fz}};
  Vec3Ptr a1{{(double*)abx1, (double*)aby1, (double*)abz1}}, a2{{(double*)abx2, (double*)aby2, (double*)abz2}};
  Vec3CPtr uu{{(const double*)u, (const double*)v, (const double*)w}};
  Vec3CPtr l1{{(const double*)ul1, (const double*)vl1, (const double*)wl1}};
  Vec3CPtr l2{{(const double*)ul2, (const double*)vl2, (const double*)wl2}};
  const int threads = 256;
  abf_bdf_kernel<<<grid_for(n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      f, a1, a2, do_abf, do_abf ? ext[0] : 0.0, do_abf ? ext[1] : 0.0, do_abf ? ext[2] : 0.0, do_bdf, uu, l1, l2,
      (const double*)B, rho, do_bdf ? rho / dt : 0.0, do_bdf ? bd[1] : 0.0, (do_bdf && nbd >= 2) ? bd[2] : 0.0,
      (do_bdf && nbd >= 3) ? bd[3] : 0.0, nbd, (int64_t)n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_sumab(void* ue, void* ve, void* we, const void* u, const void* v, const void* w, const void* ulag1,
               const void* vlag1, const void* wlag1, const void* ulag2, const void* vlag2, const void* wlag2,
               const double* ab, const int* nab, const int* n, void* stream) {
  if (!ue || !ve || !we || !u || !v || !w || !ulag1 || !vlag1 || !wlag1 || !ab || !nab || !n)
    return fail(B200_ERR_ARG, "sumab: null argument");
  if (*nab < 2 || *nab > 3) return fail(B200_ERR_ARG, "sumab: nab=%d (2 or 3)", *nab);
  if (*nab == 3 && (!ulag2 || !vlag2 || !wlag2)) return fail(B200_ERR_ARG, "sumab: nab=3 needs the second lag");
  Vec3Ptr e{{(double*)ue, (double*)ve, (double*)we}};
  Vec3CPtr uu{{(const double*)u, (const double*)v, (const double*)w}};
  Vec3CPtr l1{{(const double*)ulag1, (const double*)vlag1, (const double*)wlag1}};
  Vec3CPtr l2{{(const double*)ulag2, (const double*)vlag2, (const double*)wlag2}};
  const int threads = 256;
  sumab_kernel<<<grid_for(*n, threads, dev_sm_count(), 8), threads, 0, (cudaStream_t)stream>>>(
      e, uu, l1, l2, ab[0], ab[1], *nab == 3 ? ab[2] : 0.0, *nab, (int64_t)*n);
  LAUNCHED();
  CK(cudaGetLastError());
  return B200_OK;
}

int b200_makeabf(void* abx1, void* aby1, void* abz1, void* abx2, void* aby2, void* abz2, void* fx, void* fy,
                 void* fz, const double* rho, const double* ext, const int* n, void* stream) {
  if (!abx1 || !aby1 || !abz1 || !abx2 || !aby2 || !abz2 || !fx || !fy || !fz || !rho || !ext || !n)
    return fail(B200_ERR_ARG, "makeabf: null argument");
  return abf_bdf_launch(fx, fy, fz, abx1, aby1, abz1, abx2, aby2, abz2, 1, ext, 0, nullptr, nullptr, nullptr,
                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, *rho, 1.0, nullptr, 0, *n,
                        stream);
}

int b200_makebdf(const void* ulag1, const void* vlag1, const void* wlag1, const void* ulag2, const void* vlag2,
                 const void* wlag2, void* fx, void* fy, void* fz, const void* u, const void* v, const void* w,
                 const void* B, const double* rho, const double* dt, const double* bd, const int* nbd,
                 const int* n, void* stream) {
  if (!fx || !fy || !fz || !u || !v || !w || !B || !rho || !dt || !bd || !nbd || !n)
    return fail(B200_ERR_ARG, "makebdf: null argument");
  if (*nbd < 1 || *nbd > 3) return fail(B200_ERR_ARG, "makebdf: nbd=%d (1..3)", *nbd);
  if ((*nbd >= 2 && (!ulag1 || !vlag1 || !wlag1)) || (*nbd >= 3 && (!ulag2 || !vlag2 || !wlag2)))
    return fail(B200_ERR_ARG, "makebdf: lag fields missing for nbd=%d", *nbd);
  return abf_bdf_launch(fx, fy, fz, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 1, u, v, w,
                        ulag1, vlag1, wlag1, ulag2, vlag2, wlag2, B, *rho, *dt, bd, *nbd, *n, stream);
}

int b200_makeabf_bdf(void* abx1, void* aby1, void* abz1, void* abx2, void* aby2, void* abz2, const void* ulag1,
                     const void* vlag1, const void* wlag1, const void* ulag2, const void* vlag2,
                     const void* wlag2, void* fx, void* fy, void* fz, const void* u, const void* v,
                     const void* w, const void* B, const double* rho, const double* dt, const double* ext,
                     const double* bd, const int* nbd, const int* n, void* stream) {
  if (!abx1 || !aby1 || !abz1 || !abx2 || !aby2 || !abz2 || !fx || !fy || !fz || !u || !v || !w || !B || !rho ||
      !dt || !ext || !bd || !nbd || !n)
    return fail(B200_ERR_ARG, "makeabf_bdf: null argument");
  if (*nbd < 1 || *nbd > 3) return fail(B200_ERR_ARG, "makeabf_bdf: nbd=%d (1..3)", *nbd);
  if ((*nbd >= 2 && (!ulag1 || !vlag1 || !wlag1)) || (*nbd >= 3 && (!ulag2 || !vlag2 || !wlag2)))
    return fail(B200_ERR_ARG, "makeabf_bdf: lag fields missing for nbd=%d", *nbd);
  return abf_bdf_launch(fx, fy, fz, abx1, aby1, abz1, abx2, aby2, abz2, 1, ext, 1, u, v, w, ulag1, vlag1, wlag1,
                        ulag2, vlag2, wlag2, B, *rho, *dt, bd, *nbd, *n, stream);
}

// ---- gather-scatter set-up ---------------------------------------------------------------------
int b200_gs_init(void* handle, const int64_t* key, const int* on_device) {
  if (!handle || !key) return fail(B200_ERR_ARG, "gs_init: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n;
  cudaStream_t st = h->stream;
  cudaFree(h->gs_off); cudaFree(h->gs_dof); cudaFree(h->gs_rep);
  h->gs_off = h->gs_dof = h->gs_rep = nullptr;
  h->have_gs = false;
  free_schedule(h);
  free_xstage(h);
  free_shared(h);      // the shared-node lists index the old class list: b200_gs_init_shared must be called again
  if (n == 0) { h->nclass = 0; h->nmember = 0; h->have_gs = true; return B200_OK; }

  // every temporary lives in T: freed early where the peak matters, and on every error return
  DevTemps T;
  int64_t *d_key_in = nullptr, *d_key_own = nullptr, *d_key = nullptr, *d_comp = nullptr, *d_comp2 = nullptr;
  int *d_idx = nullptr, *d_dof = nullptr, *d_head = nullptr, *d_hscan = nullptr;
  unsigned char* d_shared = nullptr;
  unsigned char* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  if (on_device && *on_device) d_key_in = const_cast<int64_t*>(key);
  else {
    if (int r = T.alloc(&d_key_own, n)) return r;
    d_key_in = d_key_own;
    CK(cudaMemcpyAsync(d_key_in, key, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  }
  if (int r = T.alloc(&d_key, n)) return r;
  if (int r = T.alloc(&d_idx, n)) return r;
  if (int r = T.alloc(&d_dof, n)) return r;
  const int threads = 256;
  const int grid = grid_for(n, threads, h->num_sm, 8);
  gs_iota_kernel<<<grid, threads, 0, st>>>(d_idx, n);
  LAUNCHED();
  // 1. stable radix sort of (key, dof)
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key_in, d_key, d_idx, d_dof, n, 0, 64, st));
  if (int r = T.alloc(&d_tmp, tmp_bytes)) return r;
  CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key_in, d_key, d_idx, d_dof, n, 0, 64, st));
  CK(cudaStreamSynchronize(st));
  T.release(d_tmp);
  T.release(d_key_own);
  // 2. mark shared members and run heads
  if (int r = T.alloc(&d_shared, n)) return r;
  if (int r = T.alloc(&d_head, n)) return r;
  if (int r = T.alloc(&d_hscan, n)) return r;
  gs_mark_kernel<<<grid, threads, 0, st>>>(d_key, n, d_shared, d_head);
  LAUNCHED();
  CK(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, d_head, d_hscan, cub::Max(), n, st));
  if (int r = T.alloc(&d_tmp, tmp_bytes)) return r;
  CK(cub::DeviceScan::InclusiveScan(d_tmp, tmp_bytes, d_head, d_hscan, cub::Max(), n, st));
  CK(cudaStreamSynchronize(st));
  T.release(d_tmp);
  // 3. composite keys (first dof of class, dof) and rep[]
  T.release(d_key);
  if (int r = T.alloc(&d_comp, n)) return r;
  if (int r = dmalloc(&h->gs_rep, n)) return r;
  gs_compose_kernel<<<grid, threads, 0, st>>>(d_dof, d_hscan, n, d_comp, h->gs_rep);
  LAUNCHED();
  CK(cudaStreamSynchronize(st));
  T.release(d_head); T.release(d_hscan); T.release(d_idx);
  // 4. keep shared members only
  int64_t* d_ns = nullptr;
  if (int r = T.alloc(&d_ns, 1)) return r;
  if (int r = T.alloc(&d_comp2, n)) return r;
  CK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_comp, d_shared, d_comp2, d_ns, n, st));
  if (int r = T.alloc(&d_tmp, tmp_bytes)) return r;
  CK(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_comp, d_shared, d_comp2, d_ns, n, st));
  int64_t ns = 0;
  CK(cudaMemcpyAsync(&ns, d_ns, sizeof ns, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  T.release(d_tmp);
  T.release(d_shared); T.release(d_dof);
  h->nmember = ns;
  if (ns == 0) {
    h->nclass = 0;
    if (int r = dmalloc(&h->gs_off, 1)) return r;
    if (int r = dmalloc(&h->gs_dof, 1)) return r;
    h->have_gs = true;
    return B200_OK;
  }
  // 5. order members by (first dof of class, dof): element-surface order
  CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_comp2, d_comp, ns, 0, 64, st));
  if (int r = T.alloc(&d_tmp, tmp_bytes)) return r;
  CK(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_comp2, d_comp, ns, 0, 64, st));
  CK(cudaStreamSynchronize(st));
  T.release(d_tmp);
  T.release(d_comp2);
  unsigned char* d_h2 = nullptr;
  if (int r = dmalloc(&h->gs_dof, ns)) return r;
  if (int r = T.alloc(&d_h2, ns)) return r;
  const int grid2 = grid_for(ns, threads, h->num_sm, 8);
  gs_split_kernel<<<grid2, threads, 0, st>>>(d_comp, ns, h->gs_dof, d_h2);
  LAUNCHED();
  // 6. class offsets = positions of heads (+ terminator)
  int* d_off_tmp = nullptr;
  if (int r = T.alloc(&d_off_tmp, ns + 1)) return r;
  thrust::counting_iterator<int> cnt(0);
  CK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, cnt, d_h2, d_off_tmp, d_ns, ns, st));
  if (int r = T.alloc(&d_tmp, tmp_bytes)) return r;
  CK(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, cnt, d_h2, d_off_tmp, d_ns, ns, st));
  int64_t nc = 0;
  CK(cudaMemcpyAsync(&nc, d_ns, sizeof nc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->nclass = (int)nc;
  if (int r = dmalloc(&h->gs_off, nc + 1)) return r;
  CK(cudaMemcpyAsync(h->gs_off, d_off_tmp, sizeof(int) * nc, cudaMemcpyDeviceToDevice, st));
  const int ns_i = (int)ns;
  CK(cudaMemcpyAsync(h->gs_off + nc, &ns_i, sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  h->have_gs = true;
  return B200_OK;
}

int b200_gs_get_classes(void* handle, int64_t* class_id, int64_t* nclass) {
  if (!handle || !class_id) return fail(B200_ERR_ARG, "gs_get_classes: null argument");
  Handle* h = H(handle);
  if (!h->have_gs) return fail(B200_ERR_STATE, "b200_gs_init not called");
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n;
  if (n == 0) { if (nclass) *nclass = 0; return B200_OK; }
  cudaStream_t st = h->stream;
  int *d_isrep = nullptr, *d_scan = nullptr;
  int64_t* d_cid = nullptr;
  if (int r = dmalloc(&d_isrep, n)) return r;
  if (int r = dmalloc(&d_scan, n)) return r;
  if (int r = dmalloc(&d_cid, n)) return r;
  const int threads = 256;
  const int grid = grid_for(n, threads, h->num_sm, 8);
  gs_isrep_kernel<<<grid, threads, 0, st>>>(h->gs_rep, n, d_isrep);
  LAUNCHED();
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_isrep, d_scan, n, st));
  CK(cudaMalloc(&d_tmp, tmp_bytes));
  CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_isrep, d_scan, n, st));
  gs_classid_kernel<<<grid, threads, 0, st>>>(h->gs_rep, d_scan, n, d_cid);
  LAUNCHED();
  CK(cudaMemcpyAsync(class_id, d_cid, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
  int last_scan = 0, last_rep = 0;
  CK(cudaMemcpyAsync(&last_scan, d_scan + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last_rep, d_isrep + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (nclass) *nclass = (int64_t)last_scan + last_rep;
  CK(cudaFree(d_tmp)); CK(cudaFree(d_isrep)); CK(cudaFree(d_scan)); CK(cudaFree(d_cid));
  return B200_OK;
}

int b200_gs_op(void* handle, void* f) {
  if (!handle || !f) return fail(B200_ERR_ARG, "gs_op: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  NvtxRange nvtx("Velocity residual: gs_op (b200_gs_op)");
  double* f0 = (double*)f;
  if (int r = gs_launch(h, f0, f0, f0, 1)) return r;
  if (int r = gs_exchange(h, f0, f0, f0, 1)) return r;
  return gs_finish_exchange(h, f0, f0, f0, 1);
}

int b200_gs_op3(void* handle, void* fx, void* fy, void* fz) {
  if (!handle || !fx || !fy || !fz) return fail(B200_ERR_ARG, "gs_op3: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  double *f0 = (double*)fx, *f1 = (double*)fy, *f2 = (double*)fz;
  NvtxRange nvtx("Velocity residual: gs_op x3 (b200_gs_op3)");
  if (int r = gs_launch(h, f0, f1, f2, 3)) return r;
  if (int r = gs_exchange(h, f0, f1, f2, 3)) return r;
  return gs_finish_exchange(h, f0, f1, f2, 3);
}

// ---- multi-GPU ------------------------------------------------------------------------------------
int b200_comm_unique_id(char* id128) {
  if (!id128) return fail(B200_ERR_ARG, "comm_unique_id: null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NK(ncclGetUniqueId(&id));
  memcpy(id128, &id, 128);
  return B200_OK;
}

int b200_comm_init(void* handle, const char* id128, const int* rank, const int* nranks) {
  if (!handle || !id128 || !rank || !nranks) return fail(B200_ERR_ARG, "comm_init: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NK(ncclCommInitRank(&h->comm, *nranks, id, *rank));
  h->rank = *rank; h->nranks = *nranks;
  CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_recv, cudaEventDisableTiming));
  return B200_OK;
}

int b200_gs_init_shared(void* handle, const int* nshared, const int* shared_dof, const int* nneigh,
                        const int* neigh_rank, const int* neigh_off, const int* neigh_idx) {
  if (!handle || !nshared || !nneigh) return fail(B200_ERR_ARG, "gs_init_shared: null argument");
  Handle* h = H(handle);
  if (!h->have_gs) return fail(B200_ERR_STATE, "gs_init_shared: call b200_gs_init first");
  CK(cudaSetDevice(h->device));
  const int ns = *nshared, nn = *nneigh;
  free_schedule(h);
  free_xstage(h);
  free_shared(h);
  h->nshared = ns; h->nneigh = nn;
  if (ns == 0 || nn == 0) { h->nshared = 0; h->nneigh = 0; return B200_OK; }
  if (!shared_dof || !neigh_rank || !neigh_off || !neigh_idx) return fail(B200_ERR_ARG, "gs_init_shared: null list");
  // neighbours in ascending rank order so that every rank sums contributions in the same order
  std::vector<int> order(nn);
  for (int j = 0; j < nn; j++) order[j] = j;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return neigh_rank[a] < neigh_rank[b]; });
  h->neigh_rank.assign(nn, 0);
  h->neigh_off.assign(nn + 1, 0);
  std::vector<int> send_dof;
  std::vector<std::vector<std::pair<int, int>>> contrib(ns);   // (rank, src)
  for (int s = 0; s < ns; s++) contrib[s].push_back({h->rank, -1});
  for (int jj = 0; jj < nn; jj++) {
    const int j = order[jj];
    h->neigh_rank[jj] = neigh_rank[j];
    if (neigh_rank[j] == h->rank) return fail(B200_ERR_ARG, "gs_init_shared: neighbour == own rank");
    for (int i = neigh_off[j]; i < neigh_off[j + 1]; i++) {
      const int s = neigh_idx[i];
      if (s < 0 || s >= ns) return fail(B200_ERR_ARG, "gs_init_shared: neigh_idx out of range");
      contrib[s].push_back({neigh_rank[j], (int)send_dof.size()});
      send_dof.push_back(shared_dof[s]);
    }
    h->neigh_off[jj + 1] = (int)send_dof.size();
  }
  h->nsend = (int)send_dof.size();
  std::vector<int> c_off(ns + 1, 0), c_src;
  for (int s = 0; s < ns; s++) {
    std::sort(contrib[s].begin(), contrib[s].end());
    for (auto& pr : contrib[s]) c_src.push_back(pr.second);
    c_off[s + 1] = (int)c_src.size();
  }
  if (int r = dmalloc(&h->d_send_dof, send_dof.size())) return r;
  if (int r = dmalloc(&h->d_shared_dof, ns)) return r;
  if (int r = dmalloc(&h->d_s_class, ns)) return r;
  if (int r = dmalloc(&h->d_c_off, ns + 1)) return r;
  if (int r = dmalloc(&h->d_c_src, c_src.size())) return r;
  if (int r = dmalloc(&h->d_send, (size_t)h->nsend * 3)) return r;
  if (int r = dmalloc(&h->d_recv, (size_t)h->nsend * 3)) return r;
  CK(cudaMemcpy(h->d_send_dof, send_dof.data(), sizeof(int) * send_dof.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_shared_dof, shared_dof, sizeof(int) * ns, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_c_off, c_off.data(), sizeof(int) * (ns + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_c_src, c_src.data(), sizeof(int) * c_src.size(), cudaMemcpyHostToDevice));
  // local class of every shared node (-1: the node has a single local member)
  if (int r = dmalloc(&h->gs_skip, (size_t)std::max(h->nclass, 1))) return r;
  CK(cudaMemsetAsync(h->gs_skip, 0, (size_t)std::max(h->nclass, 1), h->stream));
  const int threads = 256;
  gs_find_class_kernel<<<grid_for(ns, threads, h->num_sm, 4), threads, 0, h->stream>>>(
      h->d_shared_dof, ns, h->gs_rep, h->gs_off, h->gs_dof, h->nclass, h->d_s_class, h->gs_skip);
  LAUNCHED();
  CK(cudaGetLastError());
  // compact list of local classes that hold a shared node
  std::vector<int> s_class(ns);
  CK(cudaMemcpyAsync(s_class.data(), h->d_s_class, sizeof(int) * ns, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<int> cls;
  for (int s = 0; s < ns; s++) if (s_class[s] >= 0) cls.push_back(s_class[s]);
  std::sort(cls.begin(), cls.end());
  cls.erase(std::unique(cls.begin(), cls.end()), cls.end());
  h->n_shared_cls = (int)cls.size();
  if (int r = dmalloc(&h->gs_shared_cls, cls.size())) return r;
  if (!cls.empty())
    CK(cudaMemcpy(h->gs_shared_cls, cls.data(), sizeof(int) * cls.size(), cudaMemcpyHostToDevice));
  return B200_OK;
}

// Shared-node discovery on the device + NCCL: what Neko's gs_t%init does from the dofmap
// (adjoint/adjoint_scheme.f90:339-343).  See include/neko_top_b200.h.
int b200_gs_init_shared_from_keys(void* handle, const int64_t* key, const int* on_device,
                                  const unsigned char* cand, int* nshared_out, int* nneigh_out) {
  if (!handle || !key) return fail(B200_ERR_ARG, "gs_init_shared_from_keys: null argument");
  Handle* h = H(handle);
  if (!h->have_gs) return fail(B200_ERR_STATE, "gs_init_shared_from_keys: call b200_gs_init first");
  if (!h->comm) return fail(B200_ERR_STATE, "gs_init_shared_from_keys: call b200_comm_init first");
  if (h->nranks > 64) return fail(B200_ERR_ARG, "gs_init_shared_from_keys: at most 64 ranks");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int64_t n = h->n;
  const int threads = 256, nr = h->nranks;
  const bool dev_in = on_device && *on_device;
  DevTemps T;
  const int64_t* d_key = key;
  unsigned char* d_flag = nullptr;
  if (!dev_in) {
    int64_t* tmp = nullptr;
    if (int r = T.alloc(&tmp, (size_t)n)) return r;
    CK(cudaMemcpyAsync(tmp, key, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    d_key = tmp;
  }
  if (int r = T.alloc(&d_flag, (size_t)n)) return r;
  if (cand) {
    CK(cudaMemcpyAsync(d_flag, cand, (size_t)n, dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  } else if (n > 0) {
    shk_surface_kernel<<<grid_for(n, threads, h->num_sm, 8), threads, 0, st>>>(d_flag, n, h->lx);
    LAUNCHED();
  }
  // 1. candidate dofs, their keys, stable sort by key -> unique keys with the smallest dof of each
  int *d_idx = nullptr, *d_idx2 = nullptr;
  int64_t *d_cnt1 = nullptr, *d_k = nullptr, *d_k2 = nullptr;
  void* d_tmp = nullptr;
  size_t tb = 0;
  if (int r = T.alloc(&d_idx, (size_t)n)) return r;
  if (int r = T.alloc(&d_cnt1, 2 + (size_t)nr)) return r;
  thrust::counting_iterator<int> cnt0(0);
  CK(cub::DeviceSelect::Flagged(nullptr, tb, cnt0, d_flag, d_idx, d_cnt1, n, st));
  if (int r = T.alloc(reinterpret_cast<unsigned char**>(&d_tmp), tb)) return r;
  CK(cub::DeviceSelect::Flagged(d_tmp, tb, cnt0, d_flag, d_idx, d_cnt1, n, st));
  int64_t ncand = 0;
  CK(cudaMemcpyAsync(&ncand, d_cnt1, sizeof ncand, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int m = (int)ncand;
  if (int r = T.alloc(&d_k, (size_t)m)) return r;
  if (int r = T.alloc(&d_k2, (size_t)m)) return r;
  if (int r = T.alloc(&d_idx2, (size_t)m)) return r;
  unsigned char* d_head = nullptr;
  int64_t* d_uk = nullptr;
  int* d_udof = nullptr;
  if (int r = T.alloc(&d_head, (size_t)m)) return r;
  if (int r = T.alloc(&d_uk, (size_t)m)) return r;
  if (int r = T.alloc(&d_udof, (size_t)m)) return r;
  int64_t nuk64 = 0;
  if (m > 0) {
    shk_gather_kernel<<<grid_for(m, threads, h->num_sm, 8), threads, 0, st>>>(d_key, d_idx, m, d_k);
    LAUNCHED();
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_k, d_k2, d_idx, d_idx2, m, 0, 64, st));
    void* d_tmp2 = nullptr;
    if (int r = T.alloc(reinterpret_cast<unsigned char**>(&d_tmp2), tb)) return r;
    CK(cub::DeviceRadixSort::SortPairs(d_tmp2, tb, d_k, d_k2, d_idx, d_idx2, m, 0, 64, st));
    shk_head_kernel<<<grid_for(m, threads, h->num_sm, 8), threads, 0, st>>>(d_k2, m, d_head);
    LAUNCHED();
    void* d_tmp3 = nullptr;
    CK(cub::DeviceSelect::Flagged(nullptr, tb, d_k2, d_head, d_uk, d_cnt1, m, st));
    if (int r = T.alloc(reinterpret_cast<unsigned char**>(&d_tmp3), tb)) return r;
    CK(cub::DeviceSelect::Flagged(d_tmp3, tb, d_k2, d_head, d_uk, d_cnt1, m, st));
    CK(cub::DeviceSelect::Flagged(d_tmp3, tb, d_idx2, d_head, d_udof, d_cnt1, m, st));
    CK(cudaMemcpyAsync(&nuk64, d_cnt1, sizeof nuk64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  const int nuk = (int)nuk64;
  // 2. all-gather the counts, then the padded key lists
  int64_t* d_cnts = d_cnt1 + 2;
  CK(cudaMemcpyAsync(d_cnt1, &nuk64, sizeof nuk64, cudaMemcpyHostToDevice, st));
  NK(ncclAllGather(d_cnt1, d_cnts, 1, ncclInt64, h->comm, st));
  std::vector<int64_t> cnts(nr);
  CK(cudaMemcpyAsync(cnts.data(), d_cnts, sizeof(int64_t) * nr, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int64_t mpad = 1;
  for (int r = 0; r < nr; r++) mpad = std::max(mpad, cnts[r]);
  int64_t *d_pad = nullptr, *d_all = nullptr;
  if (int r = T.alloc(&d_pad, (size_t)mpad)) return r;
  if (int r = T.alloc(&d_all, (size_t)mpad * nr)) return r;
  CK(cudaMemsetAsync(d_pad, 0xff, sizeof(int64_t) * (size_t)mpad, st));
  if (nuk > 0) CK(cudaMemcpyAsync(d_pad, d_uk, sizeof(int64_t) * (size_t)nuk, cudaMemcpyDeviceToDevice, st));
  NK(ncclAllGather(d_pad, d_all, (size_t)mpad, ncclInt64, h->comm, st));
  // 3. which of my unique candidate keys live on which other rank
  unsigned long long* d_hit = nullptr;
  if (int r = T.alloc(&d_hit, (size_t)nuk)) return r;
  if (nuk > 0) {
    shk_hit_kernel<<<grid_for(nuk, threads, h->num_sm, 8), threads, 0, st>>>(d_uk, nuk, d_all, mpad, d_cnts, nr,
                                                                            h->rank, d_hit);
    LAUNCHED();
    CK(cudaGetLastError());
  }
  std::vector<unsigned long long> hit(nuk);
  std::vector<int> udof(nuk);
  std::vector<int64_t> uk(nuk);
  std::vector<int64_t> ks(m);
  std::vector<int> kd(m);
  if (nuk > 0) {
    CK(cudaMemcpyAsync(hit.data(), d_hit, sizeof(unsigned long long) * nuk, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(udof.data(), d_udof, sizeof(int) * nuk, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(uk.data(), d_uk, sizeof(int64_t) * nuk, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ks.data(), d_k2, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(kd.data(), d_idx2, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  // 4. host: compact numbering of the shared nodes (ascending key), per-neighbour lists in ascending key order
  //    (both sides of a pair build the same order without a handshake), elements owning a shared node
  std::vector<int> sidx(nuk, -1), shared_dof;
  for (int i = 0; i < nuk; i++)
    if (hit[i]) { sidx[i] = (int)shared_dof.size(); shared_dof.push_back(udof[i]); }
  std::vector<int> neigh_rank, neigh_off(1, 0), neigh_idx;
  for (int r = 0; r < nr; r++) {
    size_t before = neigh_idx.size();
    for (int i = 0; i < nuk; i++)
      if (hit[i] >> r & 1ull) neigh_idx.push_back(sidx[i]);
    if (neigh_idx.size() > before) { neigh_rank.push_back(r); neigh_off.push_back((int)neigh_idx.size()); }
  }
  const int N3 = h->lx * h->lx * h->lx;
  std::vector<char> isb(h->nelv, 0);
  {
    int u = -1;                                   // walk the sorted candidate (key, dof) pairs and the unique keys together
    for (int i = 0; i < m; i++) {
      if (i == 0 || ks[i] != ks[i - 1]) u++;
      if (hit[u]) isb[kd[i] / N3] = 1;
    }
  }
  std::vector<int> bnd;
  for (int e = 0; e < h->nelv; e++) if (isb[e]) bnd.push_back(e);
  const int ns = (int)shared_dof.size(), nn = (int)neigh_rank.size();
  if (int r = b200_gs_init_shared(handle, &ns, shared_dof.data(), &nn, neigh_rank.data(), neigh_off.data(),
                                  neigh_idx.data()))
    return r;
  const int nb = (int)bnd.size();
  if (int r = b200_adjrhs_set_boundary_elements(handle, &nb, bnd.data())) return r;
  if (nshared_out) *nshared_out = ns;
  if (nneigh_out) *nneigh_out = nn;
  return B200_OK;
}

int b200_adjrhs_set_boundary_elements(void* handle, const int* nbnd, const int* bnd_elem) {
  if (!handle || !nbnd) return fail(B200_ERR_ARG, "set_boundary_elements: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  cudaFree(h->d_bnd_elem); cudaFree(h->d_int_elem);
  h->d_bnd_elem = h->d_int_elem = nullptr;
  h->nbnd = *nbnd; h->nint = 0;
  free_schedule(h);
  if (*nbnd == 0) return B200_OK;
  if (!bnd_elem) return fail(B200_ERR_ARG, "set_boundary_elements: null list");
  std::vector<char> isb(h->nelv, 0);
  for (int i = 0; i < *nbnd; i++) {
    if (bnd_elem[i] < 0 || bnd_elem[i] >= h->nelv) return fail(B200_ERR_ARG, "boundary element out of range");
    isb[bnd_elem[i]] = 1;
  }
  std::vector<int> bnd, inte;
  for (int i = 0; i < h->nelv; i++) {            // interior elements keep the processing order
    const int e = h->order.empty() ? i : h->order[i];
    (isb[e] ? bnd : inte).push_back(e);
  }
  h->nbnd = (int)bnd.size(); h->nint = (int)inte.size();
  if (int r = dmalloc(&h->d_bnd_elem, bnd.size())) return r;
  if (int r = dmalloc(&h->d_int_elem, inte.size())) return r;
  CK(cudaMemcpy(h->d_bnd_elem, bnd.data(), sizeof(int) * bnd.size(), cudaMemcpyHostToDevice));
  if (!inte.empty()) CK(cudaMemcpy(h->d_int_elem, inte.data(), sizeof(int) * inte.size(), cudaMemcpyHostToDevice));
  return B200_OK;
}

// ---- host-staged step (bench.py "e2e") ----------------------------------------------------------
int b200_adjrhs_step_host(void* handle, const double* vx, const double* vy, const double* vz,
                          const double* vxb, const double* vyb, const double* vzb, const double* rho,
                          double* fx, double* fy, double* fz, double* sens) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  if (!vx || !vy || !vz || !vxb || !vyb || !vzb || !rho || !fx || !fy || !fz)
    return fail(B200_ERR_ARG, "step_host: null argument");
  if (!h->have_gs) return fail(B200_ERR_STATE, "step_host: b200_gs_init not called");
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n;
  const int64_t N = (int64_t)h->lx * h->lx * h->lx;
  if (!h->stage[0]) {
    for (int i = 0; i < 11; i++) if (int r = dmalloc(&h->stage[i], n)) return r;
    CK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  }
  const double* in[7] = {vx, vy, vz, vxb, vyb, vzb, rho};
  double** st = h->stage;   // 0..6 inputs, 7..9 f, 10 sens
  // element chunks: copy chunk c+1 while chunk c computes; sens goes back as soon as its chunk is done
  const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(32, h->nelv / 256));
  while ((int)h->ev_h2d.size() < nchunk) {
    cudaEvent_t a, b;
    CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_h2d.push_back(a); h->ev_k.push_back(b);
  }
  // Single GPU, mesh order, no point-zone mask: the direct-stiffness summation is pipelined too.  After chunk
  // c the classes that COMPLETE in it are summed (they are contiguous ranges of the schedule's packed lists,
  // which are sorted by completing position), and f goes back for every leading element whose classes are
  // all summed -- so the device->host copy of f overlaps the host->device copy of the later chunks instead
  // of waiting for the whole mesh.  Same sums in the same order: bit-identical to the one-pass path.
  const bool masked = h->if_lube && h->lube_mask_size > 0;
  bool pipe_gs = !h->comm && h->order.empty() && !masked && h->nclass > 0 && nchunk > 1;
  if (pipe_gs) {
    if (!h->sched_valid || h->sched_kind != 1) {
      if (int r = build_gs_schedule(h, nullptr, h->nelv)) return r;
      h->sched_kind = 1;
    }
    pipe_gs = h->sched_valid && (int)h->sched_elem_last.size() == h->nelv;
  }
  std::vector<int> cb(nchunk + 1), ready(nchunk + 1, 0);
  for (int c = 0; c <= nchunk; c++) cb[c] = (int)((int64_t)h->nelv * c / nchunk);
  std::vector<int> eo(4 * (size_t)(nchunk + 1), 0);
  if (pipe_gs) {
    for (int c = 0; c <= nchunk; c++)
      CK(cudaMemcpyAsync(&eo[4 * c], h->sched_eoff + 4 * (size_t)cb[c], 4 * sizeof(int), cudaMemcpyDeviceToHost,
                         h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int m = 0, run_max = -1;                  // ready[c+1]: elements [0, m) are final after chunk c
    for (int c = 0; c < nchunk; c++) {
      while (m < h->nelv && std::max(run_max, h->sched_elem_last[m]) < cb[c + 1]) {
        run_max = std::max(run_max, h->sched_elem_last[m]);
        m++;
      }
      ready[c + 1] = m;
    }
  }
  // the caller's stream must be idle with respect to the staging buffers
  CK(cudaEventRecord(h->ev_done, h->stream));
  CK(cudaStreamWaitEvent(h->h2d_stream, h->ev_done, 0));
  CK(cudaStreamWaitEvent(h->d2h_stream, h->ev_done, 0));
  double* out[3] = {fx, fy, fz};
  for (int c = 0; c < nchunk; c++) {
    const int e0 = cb[c], e1 = cb[c + 1];
    const size_t o = (size_t)e0 * N, cnt = (size_t)(e1 - e0) * N;
    for (int i = 0; i < 7; i++)
      CK(cudaMemcpyAsync(st[i] + o, in[i] + o, cnt * 8, cudaMemcpyHostToDevice, h->h2d_stream));
    CK(cudaEventRecord(h->ev_h2d[c], h->h2d_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_h2d[c], 0));
    LaunchArgs a = make_args(st[0], st[1], st[2], st[3], st[4], st[5], st[6], nullptr, nullptr, nullptr,
                             nullptr, st[7], st[8], st[9], sens ? st[10] : nullptr, nullptr, e1 - e0);
    a.elem_begin = e0;
    if (int r = launch_fused(h, a)) return r;
    if (pipe_gs) {
      if (int r = gs_packed_range(h, st[7], st[8], st[9], &eo[4 * c], &eo[4 * (c + 1)])) return r;
      if (c == nchunk - 1) if (int r = gs_leftover(h, st[7], st[8], st[9])) return r;
    }
    if (sens || pipe_gs) {
      CK(cudaEventRecord(h->ev_k[c], h->stream));
      CK(cudaStreamWaitEvent(h->d2h_stream, h->ev_k[c], 0));
    }
    if (sens) CK(cudaMemcpyAsync(sens + o, st[10] + o, cnt * 8, cudaMemcpyDeviceToHost, h->d2h_stream));
    if (pipe_gs && ready[c + 1] > ready[c]) {
      const size_t fo = (size_t)ready[c] * N, fc = (size_t)(ready[c + 1] - ready[c]) * N;
      for (int i = 0; i < 3; i++)
        CK(cudaMemcpyAsync(out[i] + fo, st[7 + i] + fo, fc * 8, cudaMemcpyDeviceToHost, h->d2h_stream));
    }
  }
  if (!pipe_gs) {
    {
      LaunchArgs a = make_args(st[0], st[1], st[2], st[3], st[4], st[5], st[6], nullptr, nullptr, nullptr,
                               nullptr, st[7], st[8], st[9], nullptr, nullptr, h->nelv);
      if (int r = masked_lube_post(h, a)) return r;
    }
    if (int r = gs_launch(h, st[7], st[8], st[9], 3)) return r;
    if (int r = gs_exchange(h, st[7], st[8], st[9], 3)) return r;
    if (int r = gs_finish_exchange(h, st[7], st[8], st[9], 3)) return r;
    for (int i = 0; i < 3; i++)
      CK(cudaMemcpyAsync(out[i], st[7 + i], n * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->d2h_stream));
  CK(cudaStreamSynchronize(h->stream));
  return B200_OK;
}

// ---- diagnostics ------------------------------------------------------------------------------------
int b200_adjrhs_enable_timing(void* handle, const int* flag) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  h->timing = flag && *flag;
  h->tev_n = 0;
  if (h->timing) {   // pre-create the pool so no event is created inside a timed region
    CK(cudaSetDevice(h->device));
    while (h->tev.size() < 3 * 256) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->tev.push_back(e);
    }
  }
  return B200_OK;
}

int b200_adjrhs_get_phase_timing(void* handle, double* ms, int* nphase) {
  if (!handle || !ms || !nphase) return fail(B200_ERR_ARG, "get_phase_timing: null argument");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const int np = h->pev_n > 0 ? h->pev_n - 1 : 0;
  for (int i = 0; i < np && i < *nphase; i++) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, h->pev[i], h->pev[i + 1]));
    ms[i] = t;
  }
  *nphase = np;
  return B200_OK;
}

int b200_adjrhs_get_timing(void* handle, double* elem_kernel_ms, double* gs_ms, int64_t* launches) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  Handle* h = H(handle);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  double a = 0.0, b = 0.0;
  const size_t nt = h->tev_n / 3;
  for (size_t i = 0; i < nt; i++) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->tev[3 * i], h->tev[3 * i + 1])); a += ms;
    CK(cudaEventElapsedTime(&ms, h->tev[3 * i + 1], h->tev[3 * i + 2])); b += ms;
  }
  if (elem_kernel_ms) *elem_kernel_ms = nt ? a / nt : 0.0;
  if (gs_ms) *gs_ms = nt ? b / nt : 0.0;
  if (launches) *launches = (int64_t)nt;
  return B200_OK;
}

}  // extern "C"
