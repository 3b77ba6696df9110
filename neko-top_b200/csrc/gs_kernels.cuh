// Gather-scatter (direct-stiffness summation) for sm_100a.
// Replaces Neko's gs_t%op(., GS_OP_ADD) at /root/reference/sources/adjoint/adjoint_pnpn.f90:725,755-757.
//
// Set-up (device, CUB): radix-sort (key, dof) pairs; runs of equal keys are the node classes; classes
// with one member are dropped; the remaining members are re-sorted by (smallest dof of the class, dof)
// so that (a) members are summed in ascending dof order -- deterministic, the same order the oracle
// uses -- and (b) consecutive classes touch neighbouring addresses (element-surface order) instead of
// global-lattice order.  Result: CSR lists off[nclass+1], dof[nmember] (int32).
//
// Op: one thread per class, gather -> sum -> scatter, 1 or 3 fields per pass, no atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__global__ void gs_iota_kernel(int* idx, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) idx[i] = (int)i;
}

// after sorting by key: shared[i] = member of a class with >= 2 members; headpos[i] = i if first of run
__global__ void gs_mark_kernel(const int64_t* __restrict__ key, int64_t n, unsigned char* __restrict__ shared,
                               int* __restrict__ headpos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t k = key[i];
    const bool same_prev = (i > 0) && key[i - 1] == k;
    const bool same_next = (i + 1 < n) && key[i + 1] == k;
    shared[i] = (same_prev || same_next) ? 1 : 0;
    headpos[i] = same_prev ? 0 : (int)i;
  }
}

// composite 64-bit sort key (first dof of class << 32 | own dof), rep[dof] = first dof of its class
__global__ void gs_compose_kernel(const int* __restrict__ dof_sorted, const int* __restrict__ headpos_scan,
                                  int64_t n, int64_t* __restrict__ comp, int* __restrict__ rep) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int first = dof_sorted[headpos_scan[i]];   // stable sort => smallest dof of the class
    const int d = dof_sorted[i];
    comp[i] = ((int64_t)first << 32) | (uint32_t)d;
    rep[d] = first;
  }
}

__global__ void gs_split_kernel(const int64_t* __restrict__ comp, int64_t ns, int* __restrict__ dof,
                                unsigned char* __restrict__ head) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += stride) {
    const int64_t c = comp[i];
    dof[i] = (int)(c & 0xffffffffll);
    head[i] = (i == 0 || (comp[i - 1] >> 32) != (c >> 32)) ? 1 : 0;
  }
}

__global__ void gs_isrep_kernel(const int* __restrict__ rep, int64_t n, int* __restrict__ isrep) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    isrep[i] = (rep[i] == (int)i) ? 1 : 0;
}
__global__ void gs_classid_kernel(const int* __restrict__ rep, const int* __restrict__ scan, int64_t n,
                                  int64_t* __restrict__ cid) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cid[i] = scan[rep[i]];
}

template <int NF>
__global__ void gs_op_kernel(double* f0, double* f1, double* f2,
                             const int* __restrict__ off, const int* __restrict__ dof, int nclass) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    const int b = off[c], e = off[c + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      s0 += f0[d];
      if (NF > 1) { s1 += f1[d]; s2 += f2[d]; }
    }
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      f0[d] = s0;
      if (NF > 1) { f1[d] = s1; f2[d] = s2; }
    }
  }
}

// ---- multi-GPU shared nodes -------------------------------------------------------------------
// pack: local (already direct-stiffness-summed) value of each shared node -> send buffers, laid out
// per neighbour: buf[(off[j] + i)*NF + c]
template <int NF>
__global__ void gs_pack_kernel(const double* f0, const double* f1,
                               const double* f2, const int* __restrict__ send_dof, int nsend,
                               double* __restrict__ buf) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsend; i += stride) {
    const int d = send_dof[i];
    buf[(size_t)i * NF] = f0[d];
    if (NF > 1) { buf[(size_t)i * NF + 1] = f1[d]; buf[(size_t)i * NF + 2] = f2[d]; }
  }
}
// unpack: for every shared node, total = sum over ranks in ASCENDING RANK order (own value at its
// rank position) so every rank computes bit-identical sums; written to all local members.
// contribution lists (CSR over shared nodes): src >= 0 -> index into recv buffer; src == -1 -> own.
// s_class[s] = local class holding the node (members in the gs CSR) or -1 (single local member).
template <int NF>
__global__ void gs_unpack_kernel(double* f0, double* f1, double* f2,
                                 const double* __restrict__ recv, const int* __restrict__ c_off,
                                 const int* __restrict__ c_src, const int* __restrict__ rep_dof,
                                 const int* __restrict__ s_class, const int* __restrict__ off,
                                 const int* __restrict__ dof, int nshared) {
  const int stride = gridDim.x * blockDim.x;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nshared; s += stride) {
    const int d0 = rep_dof[s];
    const double o0 = f0[d0];
    double o1 = 0.0, o2 = 0.0;
    if (NF > 1) { o1 = f1[d0]; o2 = f2[d0]; }
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int j = c_off[s]; j < c_off[s + 1]; j++) {
      const int src = c_src[j];
      if (src < 0) { t0 += o0; t1 += o1; t2 += o2; }
      else {
        t0 += recv[(size_t)src * NF];
        if (NF > 1) { t1 += recv[(size_t)src * NF + 1]; t2 += recv[(size_t)src * NF + 2]; }
      }
    }
    const int cls = s_class[s];
    if (cls < 0) {
      f0[d0] = t0;
      if (NF > 1) { f1[d0] = t1; f2[d0] = t2; }
    } else {
      for (int m = off[cls]; m < off[cls + 1]; m++) {
        const int d = dof[m];
        f0[d] = t0;
        if (NF > 1) { f1[d] = t1; f2[d] = t2; }
      }
    }
  }
}

// local class of each shared node: classes are ordered by their first (smallest) dof, so a binary
// search over dof[off[c]] finds it; also flags the class so the bulk gs pass skips it.
__global__ void gs_find_class_kernel(const int* __restrict__ shared_dof, int nshared,
                                     const int* __restrict__ rep, const int* __restrict__ off,
                                     const int* __restrict__ dof, int nclass, int* __restrict__ s_class,
                                     unsigned char* __restrict__ skip) {
  const int stride = gridDim.x * blockDim.x;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nshared; s += stride) {
    const int r = rep[shared_dof[s]];
    int lo = 0, hi = nclass - 1, found = -1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int v = dof[off[mid]];
      if (v == r) { found = mid; break; }
      if (v < r) lo = mid + 1; else hi = mid - 1;
    }
    s_class[s] = found;
    if (found >= 0) skip[found] = 1;
  }
}

// gs over an explicit class list (the classes holding shared nodes; run before the exchange)
template <int NF>
__global__ void gs_op_list_kernel(double* f0, double* f1, double* f2,
                                  const int* __restrict__ off, const int* __restrict__ dof,
                                  const int* __restrict__ cls, int ncls) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncls; i += stride) {
    const int c = cls[i];
    const int b = off[c], e = off[c + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      s0 += f0[d];
      if (NF > 1) { s1 += f1[d]; s2 += f2[d]; }
    }
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      f0[d] = s0;
      if (NF > 1) { f1[d] = s1; f2[d] = s2; }
    }
  }
}
// gs over all classes except the flagged ones
template <int NF>
__global__ void gs_op_skip_kernel(double* f0, double* f1, double* f2,
                                  const int* __restrict__ off, const int* __restrict__ dof,
                                  const unsigned char* __restrict__ skip, int nclass) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    if (skip[c]) continue;
    const int b = off[c], e = off[c + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      s0 += f0[d];
      if (NF > 1) { s1 += f1[d]; s2 += f2[d]; }
    }
    for (int m = b; m < e; m++) {
      const int d = dof[m];
      f0[d] = s0;
      if (NF > 1) { f1[d] = s1; f2[d] = s2; }
    }
  }
}

}  // namespace b200
