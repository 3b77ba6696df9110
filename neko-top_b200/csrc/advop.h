// Internal interface between capi.cu and advop.cu (the fine-grid advection operators are a separate
// translation unit so that the two sets of heavily unrolled kernels compile in parallel).
#pragma once
#include <cuda_runtime.h>

namespace b200 {

enum : int { ADV_ADJOINT = 0, ADV_LINEAR = 1 };

struct AdvLaunch {
  int lx, lxd, mode;            // mode: ADV_ADJOINT / ADV_LINEAR (advop_kernel.cuh); lxd == lx: no dealiasing
  const double* D;              // HOST: fine-grid derivative matrix (lxd*lxd, column-major)
  const double* J;              // HOST: GLL -> fine interpolation matrix (lxd*lx, column-major); unused if lxd == lx
  const double* wd;             // HOST: fine-grid weights (lxd)
  const double* v[3];
  const double* vb[3];
  const double* G[9];           // DEVICE: geometric factors on the fine grid
  double* f[3];
  const double* rho;
  const double* B;
  const double* fs[3];
  double* sens;
  double* chi_out;
  const int* elem_list;
  int nelem;
  int elem_base;                // first element when elem_list == NULL (chunked host-buffer step)
  unsigned flags;
  double f_min, f_max, q, K_lube, K_sens;
  int num_sm;
  cudaStream_t stream;
};

// both return cudaSuccess or the failing CUDA error; *msg (static string) names an argument problem
cudaError_t advop_launch(const AdvLaunch& a, const char** msg);
cudaError_t advop_geom_to_fine(int lx, int lxd, const double* J_host, const double* const src[9],
                               double* const dst[9], int nelv, int num_sm, cudaStream_t stream,
                               const char** msg);
// strong-form derivative operators (deriv_kernels.cuh): mode DERIV_CURL_B -> out[0..2] = B*curl(u);
// DERIV_DISSIPATION -> out[0] = sum_c |grad u_c|^2
struct DerivLaunch {
  int lx, mode, nelv;
  const double* D;            // HOST, lx*lx column-major
  const double* u[3];
  const double* G[9];
  const double* jacinv;
  const double* B;
  double* out[3];
  int num_sm;
  cudaStream_t stream;
};
cudaError_t deriv_launch(const DerivLaunch& a, const char** msg);

// Helmholtz operator of the PDE filter (helm_kernels.cuh): mode 0 -> out = A u, mode 1 -> out = diag(A)
struct HelmLaunch {
  int lx, mode, nelv;
  const double* D;            // HOST
  const double* w;            // HOST
  const double* u;
  const double* G[9];
  const double* jacinv;
  const double* B;
  double* out;
  double h1, h2;
  int num_sm;
  cudaStream_t stream;
};
cudaError_t helm_launch(const HelmLaunch& a, const char** msg);

// default fine-grid order of advection_adjoint_factory (adjoint/advection_adjoint_fctry.f90:70,89): 3*lx/2
inline int advop_default_lxd(int lx) { return 3 * lx / 2; }

}  // namespace b200
