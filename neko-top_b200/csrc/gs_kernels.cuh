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

#include "adjrhs_kernel_v3.cuh"   // xs_run_begin / xs_is_run_start: the element -> slot map of the x-stage kernels

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

// UN classes per thread are in flight together (every class has >= 2 members, so the first two members of
// each are gathered unconditionally; the rare longer classes finish in a loop): the pass is bound by the
// latency of the chain offsets -> members -> values, not by bandwidth.
template <int NF, int UN = 1, int MINB = 3>
__global__ void __launch_bounds__(256, MINB) gs_op_kernel(double* f0, double* f1, double* f2,
                             const int* __restrict__ off, const int* __restrict__ dof, int nclass,
                             const unsigned char* __restrict__ skip = nullptr) {
  const int stride = gridDim.x * blockDim.x;
  for (int c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < nclass; c0 += stride * UN) {
    int b[UN], e[UN], d0[UN], d1[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int c = c0 + u * stride;
      b[u] = e[u] = 0;
      if (c < nclass && !(skip && skip[c])) { b[u] = off[c]; e[u] = off[c + 1]; }
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      d0[u] = d1[u] = 0;
      if (e[u] > b[u]) { d0[u] = dof[b[u]]; d1[u] = dof[b[u] + 1]; }
    }
    double s[UN][3];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (e[u] > b[u]) {
        const double a0 = f0[d0[u]], c0v = f0[d1[u]];
        s[u][0] = (0.0 + a0) + c0v;
        if (NF > 1) {
          const double a1 = f1[d0[u]], c1v = f1[d1[u]], a2 = f2[d0[u]], c2v = f2[d1[u]];
          s[u][1] = (0.0 + a1) + c1v;
          s[u][2] = (0.0 + a2) + c2v;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (e[u] > b[u]) {
        for (int m = b[u] + 2; m < e[u]; m++) {
          const int d = dof[m];
          s[u][0] += f0[d];
          if (NF > 1) { s[u][1] += f1[d]; s[u][2] += f2[d]; }
        }
        f0[d0[u]] = s[u][0]; f0[d1[u]] = s[u][0];
        if (NF > 1) { f1[d0[u]] = s[u][1]; f1[d1[u]] = s[u][1]; f2[d0[u]] = s[u][2]; f2[d1[u]] = s[u][2]; }
        for (int m = b[u] + 2; m < e[u]; m++) {
          const int d = dof[m];
          f0[d] = s[u][0];
          if (NF > 1) { f1[d] = s[u][1]; f2[d] = s[u][2]; }
        }
      }
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
// ---- class schedule: classes sorted by completing element position, packed by size (pipelined host step, gs mode 1) ----
// pos[e] = position of element e in the processing list (-1: not in the list == stored before the launch)
__global__ void gs_pos_kernel(const int* __restrict__ order, int norder, int* __restrict__ pos) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < norder; i += stride) pos[order ? order[i] : i] = i;
}
// sort key of a class: (bucket << 40) | completing position.  bucket 0: 2 members, 1: 3-4, 2: 5-8, 3: 9-16,
// 4: more (left to gs_op_list_kernel), 5: handled by the multi-GPU shared-node path (skip).
__global__ void gs_class_key_kernel(const int* __restrict__ off, const int* __restrict__ dof,
                                    const unsigned char* __restrict__ skip, int nclass,
                                    const int* __restrict__ pos, int npts, unsigned long long* __restrict__ key,
                                    int* __restrict__ cls) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    const int b = off[c], e = off[c + 1], deg = e - b;
    int bucket = deg <= 2 ? 0 : deg <= 4 ? 1 : deg <= 8 ? 2 : deg <= 16 ? 3 : 4;
    if (skip && skip[c]) bucket = 5;
    int pm = 0;
    for (int m = b; m < e; m++) pm = max(pm, pos[dof[m] / npts]);
    key[c] = ((unsigned long long)bucket << 40) | (unsigned long long)pm;
    cls[c] = c;
  }
}
__device__ __forceinline__ int gs_lower_bound(const unsigned long long* __restrict__ a, int n,
                                              unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// bstart[b] = first sorted index of bucket b (b = 0..6); eoff[p*4 + b] = first entry of position p in the
// list of bucket b (p = 0..nelem)
__global__ void gs_eoff_kernel(const unsigned long long* __restrict__ key, int nclass, int nelem,
                               int* __restrict__ eoff, int* __restrict__ bstart) {
  const long long total = 4ll * (nelem + 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total + 7; w += stride) {
    if (w >= total) {
      const int b = (int)(w - total);
      bstart[b] = gs_lower_bound(key, nclass, (unsigned long long)b << 40);
      continue;
    }
    const int b = (int)(w & 3);
    const long long p = w >> 2;
    const unsigned long long base = (unsigned long long)b << 40;
    eoff[w] = gs_lower_bound(key, nclass, base | (unsigned long long)p) - gs_lower_bound(key, nclass, base);
  }
}
// packed member lists, -1 padded; left[] = class ids of bucket 4
__global__ void gs_fill_kernel(const int* __restrict__ cls, const int* __restrict__ bstart,
                               const int* __restrict__ off, const int* __restrict__ dof,
                               int* __restrict__ pair, int* __restrict__ quad, int* __restrict__ oct,
                               int* __restrict__ hex, int* __restrict__ left) {
  const int stride = gridDim.x * blockDim.x;
  const int nfill = bstart[5];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nfill; i += stride) {
    const int c = cls[i];
    const int b = off[c], deg = off[c + 1] - b;
    if (i >= bstart[4]) { left[i - bstart[4]] = c; continue; }
    int* dst; int width;
    if (i < bstart[1]) { dst = pair + 2 * (size_t)(i - bstart[0]); width = 2; }
    else if (i < bstart[2]) { dst = quad + 4 * (size_t)(i - bstart[1]); width = 4; }
    else if (i < bstart[3]) { dst = oct + 8 * (size_t)(i - bstart[2]); width = 8; }
    else { dst = hex + 16 * (size_t)(i - bstart[3]); width = 16; }
    for (int m = 0; m < width; m++) dst[m] = m < deg ? dof[b + m] : -1;
  }
}

// last[e] = largest completing position among the classes that touch element e (classes with more than 16
// members count as completing at `late`): f of element e is final once the classes of all positions
// <= last[e] are summed.  last[] must be initialised with the element's own position.
__global__ void gs_elem_last_kernel(const int* __restrict__ off, const int* __restrict__ dof,
                                    const unsigned char* __restrict__ skip, int nclass,
                                    const int* __restrict__ pos, int npts, int late, int* __restrict__ last) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    if (skip && skip[c]) continue;
    const int b = off[c], e = off[c + 1];
    int pm = 0;
    for (int m = b; m < e; m++) pm = max(pm, pos[dof[m] / npts]);
    if (e - b > 16) pm = late;
    for (int m = b; m < e; m++) atomicMax(last + dof[m] / npts, pm);
  }
}

// ---- separate pass over the packed lists of the schedule -------------------------------------------------
// pairs (most classes): UN independent classes per thread -> 6*UN gathers in flight per thread; the class
// descriptor is one coalesced int2 load (no offset/member indirection as in gs_op_kernel).
template <int UN>
__global__ void __launch_bounds__(256) gs_pairs_kernel(double* __restrict__ f0, double* __restrict__ f1,
                                                      double* __restrict__ f2, const int2* __restrict__ pair,
                                                      int npair) {
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < npair; base += stride * UN) {
    int2 q[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int i = base + u * stride;
      ok[u] = i < npair;
      q[u] = ok[u] ? __ldg(pair + i) : make_int2(0, 0);
    }
    double a[UN][3], b[UN][3];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (ok[u]) {
        a[u][0] = f0[q[u].x]; b[u][0] = f0[q[u].y];
        a[u][1] = f1[q[u].x]; b[u][1] = f1[q[u].y];
        a[u][2] = f2[q[u].x]; b[u][2] = f2[q[u].y];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (ok[u]) {
        const double s0 = (0.0 + a[u][0]) + b[u][0], s1 = (0.0 + a[u][1]) + b[u][1],
                     s2 = (0.0 + a[u][2]) + b[u][2];
        f0[q[u].x] = s0; f0[q[u].y] = s0;
        f1[q[u].x] = s1; f1[q[u].y] = s1;
        f2[q[u].x] = s2; f2[q[u].y] = s2;
      }
    }
  }
}
// classes with 3..16 members: W = 4, 8 or 16 ints per class (-1 padded), members summed in list order
template <int W>
__global__ void __launch_bounds__(256) gs_wide_kernel(double* __restrict__ f0, double* __restrict__ f1,
                                                     double* __restrict__ f2, const int4* __restrict__ lst,
                                                     int ncls) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncls; i += stride) {
    int d[W];
#pragma unroll
    for (int a = 0; a < W / 4; a++) {
      const int4 q = __ldg(lst + (size_t)i * (W / 4) + a);
      d[4 * a] = q.x; d[4 * a + 1] = q.y; d[4 * a + 2] = q.z; d[4 * a + 3] = q.w;
    }
    double v[W][3];
#pragma unroll
    for (int m = 0; m < W; m++)
      if (d[m] >= 0) { v[m][0] = f0[d[m]]; v[m][1] = f1[d[m]]; v[m][2] = f2[d[m]]; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int m = 0; m < W; m++)
      if (d[m] >= 0) { s0 += v[m][0]; s1 += v[m][1]; s2 += v[m][2]; }
#pragma unroll
    for (int m = 0; m < W; m++)
      if (d[m] >= 0) { f0[d[m]] = s0; f1[d[m]] = s1; f2[d[m]] = s2; }
  }
}

// ---- staged direct-stiffness summation for lx = 8 (adjrhs_kernel_v3.cuh XS + the face passes below) -----------
// On a conforming hexahedral mesh most node classes are PRODUCTS of face pairings: a face-interior node is
// shared by the two elements glued at that face, an edge node by 2x2 and a vertex by 2x2x2 elements.  For such
// a class the sum over all members equals pair sums taken direction by direction -- x pairs, then y pairs of
// the x sums, then z pairs -- and every copy ends with the same bits (a + b is commutative).  The step uses this:
//   X  inside the element kernel: element e-1 (i = 7) and e (i = 0) when both are consecutive in one slot's run;
//   Y  gs_face_pass_kernel<1>: element A (j = 7) and B (j = 0), rows of 64 contiguous bytes;
//   Z  gs_face_pass_kernel<2>: element A (k = 7) and B (k = 0), planes of 512 contiguous bytes;
// and only the classes that are NOT such products (irregular topology, partition-boundary classes, run starts)
// stay in the class-list pass.  The face passes read and write whole sectors -- no index lists, no sector waste.
// Set-up (all on the device, verified against the class lists of b200_gs_init, nothing assumed about the mesh):
//   sg_links_kernel     face-interior 2-member classes vote for the face links of their elements;
//   sg_elem_kernel      a link exists if all 36 interior classes of the face agree on ONE neighbour element;
//   sg_succ_kernel      the inverse maps (a j = 7 / k = 7 face may be claimed by one element only);
//   sg_classify_kernel  a class of 2, 4 or 8 members is staged iff every member has a partner in exactly the
//                       same set A of directions, 2^|A| = size, the partner maps are involutions that commute and
//                       stay inside the class; its nodes then get their bit in the per-element 64-bit masks
//                       xmask (i = 0 face, bit j + 8k), ymask (j = 0 face, bit i + 8k), zmask (k = 0 face, bit i + 8j).
// dirs: bit 0 = X, 1 = Y, 2 = Z allowed (dirs = 1: only the x pairs, results bit-identical to the plain pass).
struct SgArrays {
  int* cnt[3];              // votes per element and direction (the element owning the i/j/k = 0 face)
  int* pmin[3];             // smallest / largest neighbour element voted for
  int* pmax[3];
  int* pred[3];             // neighbour across the i/j/k = 0 face (-1: no link); X: always e-1
  int* succ[3];             // neighbour across the i/j/k = 7 face
  int* scnt[3];
  unsigned long long* mask[3];
};
__device__ __forceinline__ void sg_decode(int d, int& e, int (&x)[3]) {
  e = d >> 9;
  x[0] = d & 7; x[1] = (d >> 3) & 7; x[2] = (d >> 6) & 7;
}
__device__ __forceinline__ int sg_encode(int e, const int (&x)[3]) { return (e << 9) | x[0] | (x[1] << 3) | (x[2] << 6); }

__global__ void sg_links_kernel(const int* __restrict__ off, const int* __restrict__ dof,
                                const unsigned char* __restrict__ skip, int nclass, int nelem, int nslots, int dirs,
                                SgArrays A) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    const int b = off[c];
    if (off[c + 1] - b != 2 || (skip && skip[c])) continue;
    int e0, e1, x0[3], x1[3];
    sg_decode(dof[b], e0, x0);
    sg_decode(dof[b + 1], e1, x1);
    if (e0 == e1) continue;
    for (int a = 0; a < 3; a++) {
      if (!(dirs >> a & 1)) continue;
      const int u = (a + 1) % 3, v = (a + 2) % 3;
      if (x0[u] != x1[u] || x0[v] != x1[v] || x0[u] < 1 || x0[u] > 6 || x0[v] < 1 || x0[v] > 6) continue;
      int lo = -1, hi = -1;                         // hi owns the "0" face, lo the "7" face
      if (x0[a] == 7 && x1[a] == 0) { lo = e0; hi = e1; }
      else if (x0[a] == 0 && x1[a] == 7) { lo = e1; hi = e0; }
      else continue;
      if (a == 0 && (hi != lo + 1 || xs_is_run_start(hi, nelem, nslots))) continue;   // x: consecutive in one run
      atomicAdd(A.cnt[a] + hi, 1);
      atomicMin(A.pmin[a] + hi, lo);
      atomicMax(A.pmax[a] + hi, lo);
    }
  }
}
__global__ void sg_elem_kernel(int nelem, SgArrays A) {
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride)
    for (int a = 0; a < 3; a++) {
      const bool ok = A.cnt[a][e] == 36 && A.pmin[a][e] == A.pmax[a][e];
      const int pr = ok ? A.pmin[a][e] : -1;
      A.pred[a][e] = pr;
      if (pr >= 0) { atomicAdd(A.scnt[a] + pr, 1); A.succ[a][pr] = e; }
    }
}
__global__ void sg_succ_kernel(int nelem, SgArrays A) {
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride)
    for (int a = 0; a < 3; a++) {
      const int pr = A.pred[a][e];
      if (pr >= 0 && A.scnt[a][pr] != 1) A.pred[a][e] = -1;
      if (A.scnt[a][e] != 1) A.succ[a][e] = -1;
    }
}
// partner of node (e, x) in direction a through the verified face links, or -1
__device__ __forceinline__ int sg_partner(int e, const int (&x)[3], int a, const SgArrays& A) {
  int y[3] = {x[0], x[1], x[2]};
  int pe = -1;
  if (x[a] == 0) { pe = A.pred[a][e]; y[a] = 7; }
  else if (x[a] == 7) { pe = A.succ[a][e]; y[a] = 0; }
  if (pe < 0) return -1;
  return sg_encode(pe, y);
}
__global__ void sg_classify_kernel(const int* __restrict__ off, const int* __restrict__ dof,
                                   const unsigned char* __restrict__ skip, int nclass, int dirs, SgArrays A,
                                   int* __restrict__ keep, int* __restrict__ members) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    const int b = off[c], m = off[c + 1] - b;
    bool staged = false;
    int d[8], par[8][3];
    int dirset = 0;
    if ((m == 2 || m == 4 || m == 8) && !(skip && skip[c])) {
      for (int i = 0; i < m; i++) d[i] = dof[b + i];
      staged = true;
      for (int i = 0; i < m && staged; i++) {
        int e, x[3];
        sg_decode(d[i], e, x);
        int have = 0;
        for (int a = 0; a < 3; a++) {
          par[i][a] = -1;
          if (!(dirs >> a & 1)) continue;
          const int pd = sg_partner(e, x, a, A);
          if (pd < 0) continue;
          int idx = -1;
          for (int t = 0; t < m; t++) if (d[t] == pd) idx = t;
          if (idx < 0 || idx == i) { staged = false; break; }       // the link leaves the class: not a product
          par[i][a] = idx;
          have |= 1 << a;
        }
        if (i == 0) dirset = have;
        else if (have != dirset) staged = false;
      }
      if (staged && (dirset == 0 || (1 << __popc(dirset)) != m)) staged = false;
      for (int i = 0; i < m && staged; i++)
        for (int a = 0; a < 3 && staged; a++) {
          if (!(dirset >> a & 1)) continue;
          const int pa = par[i][a];
          if (par[pa][a] != i) staged = false;                                       // involution
          for (int a2 = a + 1; a2 < 3 && staged; a2++)
            if ((dirset >> a2 & 1) && par[pa][a2] != par[par[i][a2]][a]) staged = false;   // commute
        }
      if (staged) {
        for (int i = 0; i < m; i++) {
          int e, x[3];
          sg_decode(d[i], e, x);
          for (int a = 0; a < 3; a++) {
            if (!(dirset >> a & 1) || x[a] != 0) continue;
            const int u = (a + 1) % 3, v = (a + 2) % 3;
            // bit layout: X (j + 8k), Y (i + 8k), Z (i + 8j): the lower of the two other indices first
            const int lo = (a == 0) ? x[1] : x[0], hi = (a == 2) ? x[1] : x[2];
            (void)u; (void)v;
            atomicOr(A.mask[a] + e, 1ull << (lo + 8 * hi));
          }
        }
      }
    }
    keep[c] = staged ? 0 : 1;
    members[c] = staged ? 0 : m;
  }
}

// Y (DIR = 1) / Z (DIR = 2) face pass: one warp per element B with a non-empty mask; A = pred[B].  Lane l handles
// the node pair (i = 2(l&3), +1) of row (l>>2): for Y the row index is k (rows j = 0 of B, j = 7 of A: 64
// contiguous bytes each), for Z it is j (the planes k = 0 of B, k = 7 of A: 512 contiguous bytes).
// 128-bit load that asks L2 to fetch no more than the 64 bytes it touches: a j-row is half of a 128-byte line, and
// the default fetch granularity brings the other half (the row j = 1 or 6) from DRAM as well (ncu r02i: the y
// pass read 1.59 GB for 0.80 GB of rows)
__device__ __forceinline__ double2 ld_f64x2_l2_64(const double* p) {
  double2 v;
  asm volatile("ld.global.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
template <int DIR>
__global__ void __launch_bounds__(256) gs_face_pass_kernel(double* __restrict__ f0, double* __restrict__ f1,
                                                          double* __restrict__ f2, const int* __restrict__ pred,
                                                          const unsigned long long* __restrict__ mask, int nelem) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int r = lane >> 2, i = 2 * (lane & 3);
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nelem; e += warps) {
    const unsigned long long m = mask[e];
    if (m == 0ull) continue;
    const int pe = pred[e];
    const unsigned bits = (unsigned)(m >> (i + 8 * r)) & 3u;
    if (bits == 0u) continue;
    const size_t ob = (size_t)e * 512 + (DIR == 1 ? 64 * r : 8 * r) + i;                    // (i, 0, r) or (i, r, 0)
    const size_t oa = (size_t)pe * 512 + (DIR == 1 ? 64 * r + 56 : 8 * r + 448) + i;        // (i, 7, r) or (i, r, 7)
    double* fs[3] = {f0, f1, f2};
    double2 va[3], vb[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (DIR == 1) {
        va[c] = ld_f64x2_l2_64(fs[c] + oa);
        vb[c] = ld_f64x2_l2_64(fs[c] + ob);
      } else {
        va[c] = *reinterpret_cast<const double2*>(fs[c] + oa);
        vb[c] = *reinterpret_cast<const double2*>(fs[c] + ob);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double2 s;
      s.x = va[c].x + vb[c].x;
      s.y = va[c].y + vb[c].y;
      if (bits == 3u) {
        *reinterpret_cast<double2*>(fs[c] + oa) = s;
        *reinterpret_cast<double2*>(fs[c] + ob) = s;
      } else if (bits == 1u) {
        fs[c][oa] = s.x; fs[c][ob] = s.x;
      } else {
        fs[c][oa + 1] = s.y; fs[c][ob + 1] = s.y;
      }
    }
  }
}

// compacted CSR of the kept classes (+ their shared-node skip flags)
__global__ void xs_compact_kernel(const int* __restrict__ off, const int* __restrict__ dof, int nclass,
                                  const int* __restrict__ keep, const int* __restrict__ newidx,
                                  const int* __restrict__ newoff, const unsigned char* __restrict__ skip,
                                  int* __restrict__ off2, int* __restrict__ dof2, unsigned char* __restrict__ skip2) {
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nclass; c += stride) {
    if (!keep[c]) continue;
    const int n = newidx[c], o = newoff[c], b = off[c], cnt = off[c + 1] - b;
    off2[n] = o;
    for (int m = 0; m < cnt; m++) dof2[o + m] = dof[b + m];
    if (skip2) skip2[n] = skip ? skip[c] : 0;
  }
}

// ---- shared-node discovery between ranks (b200_gs_init_shared_from_keys) ----------------------------------
// candidate flags when the caller gives no mask: every dof on the surface of its element
__global__ void shk_surface_kernel(unsigned char* __restrict__ flag, int64_t n, int lx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int N = lx * lx * lx, L = lx - 1;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n; d += stride) {
    const int l = (int)(d % N);
    const int i = l % lx, j = (l / lx) % lx, k = l / (lx * lx);
    flag[d] = (i == 0 || i == L || j == 0 || j == L || k == 0 || k == L) ? 1 : 0;
  }
}
__global__ void shk_gather_kernel(const int64_t* __restrict__ key, const int* __restrict__ idx, int m,
                                  int64_t* __restrict__ out) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) out[i] = key[idx[i]];
}
__global__ void shk_head_kernel(const int64_t* __restrict__ ks, int m, unsigned char* __restrict__ head) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    head[i] = (i == 0 || ks[i] != ks[i - 1]) ? 1 : 0;
}
// hit[i] bit r: unique candidate key i of this rank is also a candidate key of rank r (binary search in the
// all-gathered, sorted, padded key lists allk[r*mpad ..], cnt[r] valid entries)
__global__ void shk_hit_kernel(const int64_t* __restrict__ uk, int nuk, const int64_t* __restrict__ allk,
                               int64_t mpad, const int64_t* __restrict__ cnt, int nranks, int rank,
                               unsigned long long* __restrict__ hit) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nuk; i += stride) {
    const int64_t k = uk[i];
    unsigned long long h = 0;
    for (int r = 0; r < nranks; r++) {
      if (r == rank) continue;
      const int64_t* a = allk + (size_t)r * mpad;
      int64_t lo = 0, hi = cnt[r];
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < k) lo = mid + 1; else hi = mid;
      }
      if (lo < cnt[r] && a[lo] == k) h |= 1ull << r;
    }
    hit[i] = h;
  }
}

}  // namespace b200
