// Micro-benchmark (developer tool): do DFMA (vector fp64 pipe) and DMMA (fp64 tensor-core mma.sync) share
// one pipe on sm_100a, and what are their per-SM rates / dependent-issue latencies?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes tools/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode 0: DFMA only (ILP chains), 1: DMMA m8n8k4 only, 2: both interleaved in the same warp,
// 3: even warps DFMA / odd warps DMMA, 4: DMMA m16n8k8 only, 5: DFMA + m16n8k8 same warp
template <int MODE, int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double acc[ILP], c0[ILP], c1[ILP];
  double c4[ILP][4];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const double a4[4] = {a, b, a, b};
  const double b2[2] = {b, a};
#pragma unroll
  for (int i = 0; i < ILP; i++) { acc[i] = i; c0[i] = i; c1[i] = -i; for (int j = 0; j < 4; j++) c4[i][j] = i + j; }
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0 || MODE == 2 || MODE == 5 || (MODE == 3 && !(warp & 1))) acc[i] = fma(acc[i], a, b);
      if (MODE == 1 || MODE == 2 || (MODE == 3 && (warp & 1))) dmma884(c0[i], c1[i], a, b);
      if (MODE == 4 || MODE == 5) dmma1688(c4[i], a4, b2);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i] + c0[i] + c1[i] + c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ILP>
void run(const char* name, int warps, double* out, long long* cyc) {
  const int iters = 4096;
  k<MODE, ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // FMA counts per warp-iteration
  double dfma = 32.0 * ILP, d884 = 8 * 8 * 4.0 * ILP, d1688 = 16 * 8 * 8.0 * ILP;
  double per_warp = 0;
  double wd = warps, wm = warps;
  if (MODE == 3) { wd = (warps + 1) / 2; wm = warps / 2; }
  double tot = 0;
  if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5) tot += dfma * wd;
  if (MODE == 1 || MODE == 2 || MODE == 3) tot += d884 * wm;
  if (MODE == 4 || MODE == 5) tot += d1688 * wm;
  (void)per_warp;
  tot *= iters;
  printf("%-28s ILP=%d warps/SM=%2d : %8.1f FMA/clk/SM  (%lld cycles, %.3f ms, %.1f TFLOP/s)\n", name, ILP, warps,
         tot / (double)c, c, ms, tot * 148 * 2 / (ms * 1e-3) / 1e12);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16, 32}) {
    run<0, 1>("DFMA", w, out, cyc);
    run<0, 8>("DFMA", w, out, cyc);
    run<1, 1>("DMMA m8n8k4", w, out, cyc);
    run<1, 8>("DMMA m8n8k4", w, out, cyc);
    run<4, 1>("DMMA m16n8k8", w, out, cyc);
    run<4, 4>("DMMA m16n8k8", w, out, cyc);
    run<2, 8>("DFMA+DMMA884 same warp", w, out, cyc);
    run<5, 4>("DFMA+DMMA1688 same warp", w, out, cyc);
    if (w > 1) run<3, 8>("DFMA|DMMA884 alt warps", w, out, cyc);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
