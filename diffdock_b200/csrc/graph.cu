// Batched fixed-radius neighbour search (replaces torch_cluster.radius / radius_graph on the hot path:
// models/cg_model.py:477 ligand radius graph, :543-548 cross graph with per-complex cutoff, :630 bond graph).
//
// One warp per query point y_j.  The candidates are the points x_i of the same complex (segment
// [x_ptr[b], x_ptr[b+1]) of the batch-sorted x array); lanes test 32 candidates at a time and the survivors are
// ranked with ballot/popc, so the output is deterministic and sorted by (query, candidate) - i.e. already CSR-sorted
// by the convolution's target node.  Semantics follow torch_cluster's CUDA kernel: strict squared-distance test
// d^2 < r^2, at most max_neighbors hits per query, the first ones in candidate order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diffdock_b200.h"

namespace {

template <bool FILL>
__global__ void radius_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ x_ptr,
                              const int* __restrict__ y_batch, const float* __restrict__ r_per_graph, float r_scalar,
                              int n_y, int max_neighbors, int exclude_self, int* __restrict__ count,
                              const int* __restrict__ row_start, int* __restrict__ out_row, int* __restrict__ out_col) {
  const int lane = threadIdx.x & 31;
  const int q = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  if (q >= n_y) return;
  const int b = y_batch[q];
  // per-complex cutoff c: the reference calls radius(x / c, y / c, r=1) (models/cg_model.py:543-545); the same
  // arithmetic is used here (IEEE division of every coordinate, unfused multiply/add) so that borderline pairs
  // fall on the same side of the strict test as in the reference formulation.
  const float c = r_per_graph ? r_per_graph[b] : 1.0f;
  const float r2 = __fmul_rn(r_scalar, r_scalar);
  const float qx = __fdiv_rn(y[3 * q], c), qy = __fdiv_rn(y[3 * q + 1], c), qz = __fdiv_rn(y[3 * q + 2], c);
  const int beg = x_ptr[b], end = x_ptr[b + 1];
  int found = 0;      // hits so far, including a skipped self hit (torch_cluster counts it against the cap)
  int written = 0;
  const int base = FILL ? row_start[q] : 0;
  for (int i0 = beg; i0 < end && found < max_neighbors; i0 += 32) {
    const int i = i0 + lane;
    bool hit = false;
    if (i < end) {
      const float dx = __fsub_rn(__fdiv_rn(x[3 * i], c), qx), dy = __fsub_rn(__fdiv_rn(x[3 * i + 1], c), qy),
                  dz = __fsub_rn(__fdiv_rn(x[3 * i + 2], c), qz);
      hit = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < r2;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int before = __popc(m & ((1u << lane) - 1u));
    const bool keep = hit && (found + before) < max_neighbors;
    const bool emit = keep && !(exclude_self && i == q);
    const unsigned em = __ballot_sync(0xffffffffu, emit);
    if (FILL && emit) {
      const int pos = base + written + __popc(em & ((1u << lane) - 1u));
      out_row[pos] = q;
      out_col[pos] = i;
    }
    written += __popc(em);
    found += __popc(m);
  }
  if (!FILL && lane == 0) count[q] = written;
}

}  // namespace

extern "C" {

// Pass 1: count[j] = number of neighbours of y_j (after cap / self exclusion).
int ddb200_radius_count(const float* x, const float* y, const int32_t* x_ptr, const int32_t* y_batch,
                        const float* r_per_graph, float r, int64_t n_y, int max_neighbors, int exclude_self,
                        int32_t* count, void* stream) {
  if (!x || !y || !x_ptr || !y_batch || !count || n_y < 0 || max_neighbors <= 0) return DDB200_EINVAL;
  if (n_y == 0) return 0;
  const int threads = 256;
  const long long blocks = (n_y * 32 + threads - 1) / threads;
  radius_kernel<false><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      x, y, x_ptr, y_batch, r_per_graph, r, (int)n_y, max_neighbors, exclude_self, count, nullptr, nullptr, nullptr);
  return (int)cudaGetLastError();
}

// Pass 2: row_start = exclusive scan of count; writes out_row[e] = query index, out_col[e] = candidate index.
int ddb200_radius_fill(const float* x, const float* y, const int32_t* x_ptr, const int32_t* y_batch,
                       const float* r_per_graph, float r, int64_t n_y, int max_neighbors, int exclude_self,
                       const int32_t* row_start, int32_t* out_row, int32_t* out_col, void* stream) {
  if (!x || !y || !x_ptr || !y_batch || !row_start || !out_row || !out_col || n_y < 0 || max_neighbors <= 0)
    return DDB200_EINVAL;
  if (n_y == 0) return 0;
  const int threads = 256;
  const long long blocks = (n_y * 32 + threads - 1) / threads;
  radius_kernel<true><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      x, y, x_ptr, y_batch, r_per_graph, r, (int)n_y, max_neighbors, exclude_self, nullptr, row_start, out_row, out_col);
  return (int)cudaGetLastError();
}

}  // extern "C"
