// Input side: the receptor contact graph on the device.
//
// Replaces the cdist + Python loop of datasets/process_mols.py:168-192 (residues) and :205-224 (atoms):
//     distances = cdist(coords, coords)                                  # torch.cdist, fp32
//     for i: dst = where(distances[i, :] < cutoff) minus {i}
//            if len(dst) > max_neighbors: dst = argsort(distances[i, :])[1 : max_neighbors + 1]
//            if len(dst) == 0:            dst = argsort(distances[i, :])[1 : 2]
//     edge_index = [dst_list, src_list]                                  # [neighbour, centre], listed centre by centre
// which costs O(N^2) host memory and seconds (residues) to minutes (atoms) in the reference.
//
// One warp per centre i.  Pass 1 scans all points, keeps the hits (d < cutoff, j != i) of the centre in a per-warp
// shared-memory list and counts them; the output of a centre is
//     count <= K : the hits in ascending index order                     (np.where order)
//     count >  K : the K nearest points by (distance, index)             (argsort order; ties - which np.argsort's
//                                                                         introsort leaves unspecified - by index)
//     count == 0 : the nearest other point
// Two launches around the caller's exclusive scan of the counts (same protocol as ddb200_radius_count / _fill).
//
// Distances follow torch.cdist's arithmetic, because membership and order at the cut-off / at rank K depend on their
// rounding: for more than 25 points ATen's _euclidean_dist forms  [-2 x_i, |x_i|^2, 1] . [x_j, 1, |x_j|^2]  with an
// sgemm over K = 5 - one FMA chain in k order (checked bit for bit against torch 2.11 / MKL in tests/test_inputs_cpu.py) -
// then clamp_min(0); |x|^2 = (x*x + y*y) + z*z with separately rounded products.  For <= 25 points torch uses the direct
// form sum (a - b)^2, evaluated here as dx*dx -> fma(dy, dy, .) -> fma(dz, dz, .).  The square root is the correctly
// rounded one: torch's vectorised CPU sqrt is 1 ulp off for ~0.7 % of its arguments on the AVX-512 build measured - a
// property of the host build that can only move a pair lying within one ulp of the cut-off, and is not copied.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diffdock_b200.h"

namespace {

constexpr int WARPS = 4;
constexpr int CAP = 1024;          // hits kept per centre in shared memory; beyond that the selection rescans global memory

__device__ __forceinline__ float norm2(const float* __restrict__ p) {
  return __fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2]));
}
// torch.cdist(coords, coords)[i, j] in fp32 (see the header comment); `mm` selects the matrix-multiply form
__device__ __forceinline__ float cdist_ij(const float* __restrict__ pos, int i, int j, float ni, bool mm) {
  const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  const float xj = pos[3 * j], yj = pos[3 * j + 1], zj = pos[3 * j + 2];
  if (mm) {
    const float nj = norm2(pos + 3 * j);
    float acc = __fmul_rn(-2.0f * xi, xj);
    acc = __fmaf_rn(-2.0f * yi, yj, acc);
    acc = __fmaf_rn(-2.0f * zi, zj, acc);
    acc = __fadd_rn(ni, acc);          // fma(ni, 1, acc)
    acc = __fadd_rn(nj, acc);          // fma(1, nj, acc)
    return __fsqrt_rn(fmaxf(acc, 0.f));
  }
  const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// smallest (d, j) pair strictly greater than (d_prev, j_prev) over the warp; j = -1 if none
__device__ __forceinline__ void warp_min_pair(float& d, int& j) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
    const int j2 = __shfl_xor_sync(0xffffffffu, j, o);
    if (j2 >= 0 && (j < 0 || d2 < d || (d2 == d && j2 < j))) { d = d2; j = j2; }
  }
}

template <bool FILL>
__global__ void __launch_bounds__(WARPS * 32) contact_kernel(const float* __restrict__ pos, int n, float cutoff, int K,
                                                             int knn_only, int* __restrict__ count,
                                                             const int* __restrict__ row_start, int* __restrict__ out_nbr,
                                                             int* __restrict__ out_ctr) {
  __shared__ float sD[WARPS][CAP];
  __shared__ int sJ[WARPS][CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS + warp;
  if (i >= n) return;
  const bool mm = n > 25;
  const float ni = norm2(pos + 3 * i);
  // ---- pass 1: hits in index order -----------------------------------------------------------------------------------
  int hits = 0;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    float d = 0.f;
    bool hit = false;
    if (j < n && j != i) {
      d = cdist_ij(pos, i, j, ni, mm);
      hit = knn_only ? true : d < cutoff;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int slot = hits + __popc(m & ((1u << lane) - 1u));
      if (slot < CAP) { sD[warp][slot] = d; sJ[warp][slot] = j; }
    }
    hits += __popc(m);
  }
  __syncwarp();
  const int n_out = hits == 0 ? (n > 1 ? 1 : 0) : (hits < K ? hits : K);
  if (!FILL) {
    if (lane == 0) count[i] = n_out;
    return;
  }
  const int base = row_start[i];
  if (!knn_only && hits > 0 && hits <= K) {          // np.where order
    if (hits <= CAP) {
      for (int s = lane; s < hits; s += 32) { out_nbr[base + s] = sJ[warp][s]; out_ctr[base + s] = i; }
    } else {                                         // K >= hits > CAP (huge cut-offs): re-scan and rank
      int w = 0;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = j < n && j != i && cdist_ij(pos, i, j, ni, mm) < cutoff;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int s = w + __popc(m & ((1u << lane) - 1u)); out_nbr[base + s] = j; out_ctr[base + s] = i; }
        w += __popc(m);
      }
    }
    return;
  }
  // ---- selection: the n_out nearest by (distance, index), ascending -----------------------------------------------------
  // candidates: the hit list when it is complete, else (no hit at all, or more hits than the list holds) all other points
  const bool from_list = hits > 0 && hits <= CAP;
  float d_prev = -1.f;
  int j_prev = -1;
  for (int r = 0; r < n_out; ++r) {
    float bd = 0.f;
    int bj = -1;
    if (from_list) {
      for (int s = lane; s < hits; s += 32) {
        const float d = sD[warp][s];
        const int j = sJ[warp][s];
        const bool after = d > d_prev || (d == d_prev && j > j_prev);
        if (after && (bj < 0 || d < bd || (d == bd && j < bj))) { bd = d; bj = j; }
      }
    } else {
      for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const float d = cdist_ij(pos, i, j, ni, mm);
        if (hits > 0 && !knn_only && !(d < cutoff)) continue;
        const bool after = d > d_prev || (d == d_prev && j > j_prev);
        if (after && (bj < 0 || d < bd || (d == bd && j < bj))) { bd = d; bj = j; }
      }
    }
    warp_min_pair(bd, bj);
    if (lane == 0) { out_nbr[base + r] = bj; out_ctr[base + r] = i; }
    d_prev = bd; j_prev = bj;
  }
}

}  // namespace

extern "C" int ddb200_contact_count(const float* pos, int32_t n, float cutoff, int32_t max_neighbors, int32_t knn_only,
                                    int32_t* count, void* stream) {
  if (!pos || !count || n < 0 || max_neighbors <= 0) return DDB200_EINVAL;
  if (n == 0) return 0;
  contact_kernel<false><<<(n + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream>>>(pos, n, cutoff, max_neighbors,
                                                                                          knn_only, count, nullptr, nullptr,
                                                                                          nullptr);
  return (int)cudaGetLastError();
}

extern "C" int ddb200_contact_fill(const float* pos, int32_t n, float cutoff, int32_t max_neighbors, int32_t knn_only,
                                   const int32_t* row_start, int32_t* out_nbr, int32_t* out_ctr, void* stream) {
  if (!pos || !row_start || !out_nbr || !out_ctr || n < 0 || max_neighbors <= 0) return DDB200_EINVAL;
  if (n == 0) return 0;
  contact_kernel<true><<<(n + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream>>>(pos, n, cutoff, max_neighbors,
                                                                                         knn_only, nullptr, row_start,
                                                                                         out_nbr, out_ctr);
  return (int)cudaGetLastError();
}
