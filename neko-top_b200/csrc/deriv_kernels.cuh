// Strong-form derivative operators of the minimum-dissipation objective chain, fp64.
//
// Reference (relative to /root/reference/sources; the operators themselves are Neko's, restated):
//   source_terms/adjoint_minimum_dissipation_source_term.f90:231-243   f += obj_scale * curl(curl(u)) [masked]
//   objectives/minimum_dissipation_objective_function.f90:200-222     sum_c |grad u_c|^2 integrated with B
// Neko's curl = strong derivatives (dudxyz: du = jacinv*(dr*ur + ds*us + dt*ut)), then w *= B, gs_op(ADD),
// w *= Binv.  The reference runs 6 dudxyz + 3 sub3 + opcolv sweeps per curl (9 for the objective's three
// grad calls + 9 col3/addcol3); here one element kernel produces B*curl(u) (or the objective density) from
// the three velocity components staged once in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum : int { DERIV_CURL_B = 0, DERIV_DISSIPATION = 1 };

template <int LX>
struct DerivParams {
  double D[LX * LX];       // D(i,m) at D[i + LX*m]
  const double* u[3];
  const double* G[9];      // drdx,dsdx,dtdx, drdy,dsdy,dtdy, drdz,dsdz,dtdz
  const double* jacinv;
  const double* B;         // DERIV_CURL_B: result is multiplied by B (Neko curl: opcolv before gs)
  double* out[3];          // DERIV_CURL_B: w1,w2,w3 ; DERIV_DISSIPATION: out[0] = sum_c |grad u_c|^2
  int nelv;
};

template <int LX, int MODE>
__global__ void __launch_bounds__(((LX * LX + 31) / 32) * 32)
deriv_kernel(const __grid_constant__ DerivParams<LX> p) {
  constexpr int N = LX * LX * LX, PL = LX * LX, NTHR = ((PL + 31) / 32) * 32;
  __shared__ double U[3][N];
  __shared__ double Ds[PL];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < PL; idx += NTHR) Ds[idx] = p.D[idx];
  const bool act = tid < PL;
  const int i = act ? tid % LX : 0, j = act ? tid / LX : 0;
  for (int e = blockIdx.x; e < p.nelv; e += gridDim.x) {
    const size_t eb = (size_t)e * N;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; c++)
      for (int idx = tid; idx < N; idx += NTHR) U[c][idx] = __ldg(p.u[c] + eb + idx);
    __syncthreads();
    if (!act) continue;
#pragma unroll
    for (int k = 0; k < LX; k++) {
      const int pidx = tid + PL * k;
      double g[9];
#pragma unroll
      for (int a = 0; a < 9; a++) g[a] = __ldg(p.G[a] + eb + pidx);
      const double ji = __ldg(p.jacinv + eb + pidx);
      double dx[3], dy[3], dz[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double r = 0.0, s = 0.0, t = 0.0;
#pragma unroll
        for (int m = 0; m < LX; m++) {
          r = fma(Ds[i + LX * m], U[c][m + LX * j + PL * k], r);
          s = fma(Ds[j + LX * m], U[c][i + LX * m + PL * k], s);
          t = fma(p.D[k + LX * m], U[c][tid + PL * m], t);
        }
        dx[c] = ji * (g[0] * r + g[1] * s + g[2] * t);
        dy[c] = ji * (g[3] * r + g[4] * s + g[5] * t);
        dz[c] = ji * (g[6] * r + g[7] * s + g[8] * t);
      }
      if constexpr (MODE == DERIV_CURL_B) {
        const double b = __ldg(p.B + eb + pidx);
        p.out[0][eb + pidx] = (dy[2] - dz[1]) * b;     // dw/dy - dv/dz
        p.out[1][eb + pidx] = (dz[0] - dx[2]) * b;     // du/dz - dw/dx
        p.out[2][eb + pidx] = (dx[1] - dy[0]) * b;     // dv/dx - du/dy
      } else {
        double o = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) { o += dx[c] * dx[c]; o += dy[c] * dy[c]; o += dz[c] * dz[c]; }
        p.out[0][eb + pidx] = o;
      }
    }
  }
}

// deterministic dot product: per-block partial sums of a_i*b_i over all i (mask == nullptr) or over the
// 1-based mask indices; the host adds the partials in order.  (glsc2 / glsc2_mask, local part)
static __global__ void __launch_bounds__(256) dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                         const int* __restrict__ mask, int64_t count,
                                                         double* __restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = (lo + per < count) ? lo + per : count;
  for (int64_t q = lo + threadIdx.x; q < hi; q += 256) {
    const int64_t idx = mask ? (int64_t)mask[q] - 1 : q;
    s += a[idx] * b[idx];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// out = (u + v + w) * chi   (minimum_dissipation_objective_function.f90:230-232 as written)
static __global__ void lube_density_kernel(double* __restrict__ out, const double* __restrict__ u,
                                    const double* __restrict__ v, const double* __restrict__ w,
                                    const double* __restrict__ chi, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double o = u[i] * chi[i];
    o += v[i] * chi[i];
    o += w[i] * chi[i];
    out[i] = o;
  }
}

// f_c += s * w_c on all points (mask == nullptr) or on the 1-based mask indices
// (mask_exterior_const(w, mask, 0) followed by field_add2s2)
static __global__ void add2s2_mask3_kernel(double* __restrict__ f0, double* __restrict__ f1, double* __restrict__ f2,
                                    const double* __restrict__ w0, const double* __restrict__ w1,
                                    const double* __restrict__ w2, double s, const int* __restrict__ mask,
                                    int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
    const int64_t i = mask ? (int64_t)mask[q] - 1 : q;
    f0[i] = f0[i] + s * w0[i];
    f1[i] = f1[i] + s * w1[i];
    f2[i] = f2[i] + s * w2[i];
  }
}

// mask_ops.f90:55-82 mask_exterior_const: work = c everywhere; work[mask] = fld[mask]; fld = work
static __global__ void fill_kernel(double* __restrict__ a, double c, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = c;
}
static __global__ void copy_mask_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                 const int* __restrict__ mask, int mask_size) {
  const int stride = gridDim.x * blockDim.x;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < mask_size; m += stride) {
    const int64_t i = (int64_t)mask[m] - 1;
    dst[i] = src[i];
  }
}

}  // namespace b200
