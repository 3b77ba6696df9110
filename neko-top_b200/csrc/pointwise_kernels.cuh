// Un-fused point-wise drop-ins, one per reference plug-in method (include/neko_top_b200.h).
// Pure streaming kernels: 128-bit coalesced loads/stores, grid sized as a multiple of the SM count,
// grid-stride loops.  Citations relative to /root/reference/sources.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__device__ __forceinline__ double2 ld2(const double* p, int64_t i) {
  return *reinterpret_cast<const double2*>(p + i);
}
__device__ __forceinline__ void st2(double* p, int64_t i, double2 v) {
  *reinterpret_cast<double2*>(p + i) = v;
}

// simple_brinkman_source_term.f90:149-151  field_subcol3(f, u, chi): f = f - u*chi
__global__ void brinkman_kernel(double* __restrict__ fu, double* __restrict__ fv, double* __restrict__ fw,
                                const double* __restrict__ u, const double* __restrict__ v,
                                const double* __restrict__ w, const double* __restrict__ chi, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  for (; i + 1 < n; i += stride) {
    const double2 c = ld2(chi, i);
    double2 a = ld2(fu, i), b = ld2(u, i);
    a.x -= b.x * c.x; a.y -= b.y * c.y; st2(fu, i, a);
    a = ld2(fv, i); b = ld2(v, i);
    a.x -= b.x * c.x; a.y -= b.y * c.y; st2(fv, i, a);
    a = ld2(fw, i); b = ld2(w, i);
    a.x -= b.x * c.x; a.y -= b.y * c.y; st2(fw, i, a);
  }
  if (i < n) {  // odd tail
    fu[i] -= u[i] * chi[i]; fv[i] -= v[i] * chi[i]; fw[i] -= w[i] * chi[i];
  }
}

// adjoint_lube_source_term.f90:189-203  f_i += u_i*(chi*K), everywhere
__global__ void lube_kernel(double* __restrict__ fu, double* __restrict__ fv, double* __restrict__ fw,
                            const double* __restrict__ u, const double* __restrict__ v,
                            const double* __restrict__ w, const double* __restrict__ chi, double K,
                            int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double c = chi[i] * K;
    fu[i] += u[i] * c; fv[i] += v[i] * c; fw[i] += w[i] * c;
  }
}
// same, restricted to a point zone: mask holds 1-based indices (mask_ops.f90:55-82 zeroes the rest)
__global__ void lube_mask_kernel(double* __restrict__ fu, double* __restrict__ fv, double* __restrict__ fw,
                                 const double* __restrict__ u, const double* __restrict__ v,
                                 const double* __restrict__ w, const double* __restrict__ chi, double K,
                                 const int* __restrict__ mask, int mask_size) {
  const int stride = gridDim.x * blockDim.x;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < mask_size; m += stride) {
    const int64_t i = (int64_t)mask[m] - 1;
    const double c = chi[i] * K;
    fu[i] += u[i] * c; fv[i] += v[i] * c; fw[i] += w[i] * c;
  }
}
// masked lube contribution added after the fused kernel: f_i += B*(K*RAMP(rho))*vb_i on the mask
__global__ void lube_mask_post_kernel(double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                                      const double* __restrict__ ub, const double* __restrict__ vb,
                                      const double* __restrict__ wb, const double* __restrict__ rho_or_chi,
                                      const double* __restrict__ B, double K, int do_ramp, int convex_up,
                                      double f_min, double f_max, double q,
                                      const int* __restrict__ mask, int mask_size) {
  const int stride = gridDim.x * blockDim.x;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < mask_size; m += stride) {
    const int64_t i = (int64_t)mask[m] - 1;
    double chi = rho_or_chi[i];
    if (do_ramp) {
      if (convex_up) chi = f_min + (f_max - f_min) * chi * (1.0 + q) / (chi + q);
      else chi = f_min + (f_max - f_min) * chi / (1.0 + q * (1.0 - chi));
    }
    const double c = (chi * K) * B[i];
    fx[i] += ub[i] * c; fy[i] += vb[i] * c; fz[i] += wb[i] * c;
  }
}

// adjoint_pnpn.f90:672-676 opcolv
__global__ void opcolv_kernel(double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                              const double* __restrict__ B, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double b = B[i];
    fx[i] *= b; fy[i] *= b; fz[i] *= b;
  }
}

// RAMP_mapping.f90:182-196 / 227-241
__global__ void ramp_kernel(double* __restrict__ chi, const double* __restrict__ rho, int64_t n, double f_min,
                            double f_max, double q, int convex_up) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double r = rho[i];
    chi[i] = convex_up ? f_min + (f_max - f_min) * r * (1.0 + q) / (r + q)
                       : f_min + (f_max - f_min) * r / (1.0 + q * (1.0 - r));
  }
}
// RAMP_mapping.f90:203-222 / 248-267
__global__ void ramp_backward_kernel(double* __restrict__ out, const double* __restrict__ dF,
                                     const double* __restrict__ rho, int64_t n, double f_min, double f_max,
                                     double q, int convex_up) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double r = rho[i];
    if (convex_up) {
      out[i] = (f_max - f_min) * (q + 1.0) / ((r + q) * (r + q)) * dF[i];
    } else {
      const double d = 1.0 - q * (r - 1.0);
      out[i] = (f_max - f_min) * (q + 1.0) / (d * d) * dF[i];
    }
  }
}

// minimum_dissipation_objective_function.f90:260-301
__global__ void sensitivity_kernel(double* __restrict__ S, const double* __restrict__ u,
                                   const double* __restrict__ v, const double* __restrict__ w,
                                   const double* __restrict__ ua, const double* __restrict__ va,
                                   const double* __restrict__ wa, double K, int if_lube, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double s = u[i] * ua[i];
    s += v[i] * va[i];
    s += w[i] * wa[i];
    s *= -1.0;
    if (if_lube) {
      double l = u[i] * u[i];
      l += v[i] * v[i];
      l += w[i] * w[i];
      s += K * l;
    }
    S[i] = s;
  }
}

// one-time repack of coef_t's geometry (9 cofactors + B, separate arrays x(lx,lx,lx,nelv)) into the fused
// kernel's per-plane interleaved image out[((e*lx + k)*10 + a)*lx*lx + p]  (p = i + lx*j)
struct GeomPtrs { const double* p[10]; };
__global__ void geom_pack_kernel(GeomPtrs g, double* __restrict__ out, int plane, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t pl = i / plane;            // global plane index e*lx + k
    const int64_t pt = i - pl * plane;
#pragma unroll
    for (int a = 0; a < 10; a++) out[(pl * 10 + a) * plane + pt] = g.p[a][i];
  }
}

// steady_simcomp.f90:158-176 for one field: d = old - new; acc += d*d; old = new.  Deterministic: every CTA of a
// FIXED grid (STEADY_BLOCKS, independent of n and of the device) writes one partial sum, steady_reduce_kernel adds
// the partials in a fixed tree -- the same bits on every run (field_glsc2 on one rank is a plain ordered sum too).
constexpr int STEADY_BLOCKS = 1024;
__global__ void __launch_bounds__(256) steady_update_kernel(double* __restrict__ partial, const double* __restrict__ x,
                                                            double* __restrict__ x_old, int64_t n) {
  __shared__ double red[8];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double xn = x[i];
    const double d = x_old[i] - xn;
    acc = fma(d, d, acc);
    x_old[i] = xn;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    partial[blockIdx.x] = t;
  }
}
// one CTA of 256 threads: result = sum of the STEADY_BLOCKS partials (fixed order)
__global__ void __launch_bounds__(256) steady_reduce_kernel(const double* __restrict__ partial, double* __restrict__ result) {
  __shared__ double red[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < STEADY_BLOCKS; i += 256) t += partial[i];
  red[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *result = red[0];
}

// ---- explicit time scheme around the RHS (adjoint_pnpn.f90:665-666,688-696; Neko rhs_maker types) --------
struct Vec3Ptr { double* p[3]; };
struct Vec3CPtr { const double* p[3]; };

// sumab%compute_fluid: u_e = ab1*u + ab2*ulag1 [+ ab3*ulag2]
__global__ void sumab_kernel(Vec3Ptr ue, Vec3CPtr u, Vec3CPtr l1, Vec3CPtr l2, double ab1, double ab2, double ab3,
                             int nab, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double r = ab1 * u.p[c][i] + ab2 * l1.p[c][i];
      if (nab == 3) r = r + ab3 * l2.p[c][i];
      ue.p[c][i] = r;
    }
  }
}
// makeabf%compute_fluid [+ makebdf%compute_fluid when do_bdf]: one pass over f instead of two
//   ta = ext2*f_lag + ext3*f_laglag ; f_laglag = f_lag ; f_lag = f ; f = (ext1*f + ta)*rho
//   tb = u*B*bd2 [+ ulag1*B*bd3] [+ ulag2*B*bd4] ; f = f + tb*(rho/dt)
__global__ void abf_bdf_kernel(Vec3Ptr f, Vec3Ptr ab1, Vec3Ptr ab2, int do_abf, double ext1, double ext2,
                               double ext3, int do_bdf, Vec3CPtr u, Vec3CPtr l1, Vec3CPtr l2,
                               const double* __restrict__ B, double rho, double rho_dt, double bd2, double bd3,
                               double bd4, int nbd, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double b = do_bdf ? B[i] : 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double fv = f.p[c][i];
      if (do_abf) {
        const double a1 = ab1.p[c][i], a2 = ab2.p[c][i];
        const double ta = ext2 * a1 + ext3 * a2;
        ab2.p[c][i] = a1;
        ab1.p[c][i] = fv;
        fv = (ext1 * fv + ta) * rho;
      }
      if (do_bdf) {
        double tb = u.p[c][i] * b * bd2;
        if (nbd >= 2) tb = tb + l1.p[c][i] * b * bd3;
        if (nbd >= 3) tb = tb + l2.p[c][i] * b * bd4;
        fv = fv + tb * rho_dt;
      }
      f.p[c][i] = fv;
    }
  }
}

}  // namespace b200
