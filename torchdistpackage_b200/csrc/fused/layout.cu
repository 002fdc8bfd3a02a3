// Layout kernels around attention: strided 3-D row copies with 16-byte vectors so that the
// [B,H,T,D] <-> [B,T,H,D] (and packed qkv gradient) permutes run at HBM speed instead of through
// generic strided elementwise / cat kernels.
#include <cuda_bf16.h>
#include "../common/ptx.cuh"
#include "../common/tdp_api.h"

namespace tdp {

namespace {

// copy rows of `vec_per_row` 16-byte vectors; row index = (i0, i1, i2) over (n0, n1, n2);
// strides are in 16-byte units.
__global__ void __launch_bounds__(256)
rows_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n0, int n1, int n2,
                 int vec_per_row, long s0, long s1, long s2, long d0, long d1, long d2) {
  const long total = static_cast<long>(n0) * n1 * n2 * vec_per_row;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = static_cast<int>(i % vec_per_row);
    long r = i / vec_per_row;
    const int i2 = static_cast<int>(r % n2); r /= n2;
    const int i1 = static_cast<int>(r % n1);
    const int i0 = static_cast<int>(r / n1);
    dst[i0 * d0 + i1 * d1 + i2 * d2 + v] = src[i0 * s0 + i1 * s1 + i2 * s2 + v];
  }
}

// three sources with identical strides into three destination windows of one tensor (dq | dk | dv
// -> packed dqkv): blockIdx.y selects the source, one launch instead of three
struct Rows3 {
  const uint4* src[3];
  uint4* dst[3];
};
__global__ void __launch_bounds__(256)
rows_copy3_kernel(Rows3 p, int n0, int n1, int n2, int vec_per_row, long s0, long s1, long s2,
                  long d0, long d1, long d2) {
  const uint4* __restrict__ src = p.src[blockIdx.y];
  uint4* __restrict__ dst = p.dst[blockIdx.y];
  const long total = static_cast<long>(n0) * n1 * n2 * vec_per_row;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = static_cast<int>(i % vec_per_row);
    long r = i / vec_per_row;
    const int i2 = static_cast<int>(r % n2); r /= n2;
    const int i1 = static_cast<int>(r % n1);
    const int i0 = static_cast<int>(r / n1);
    dst[i0 * d0 + i1 * d1 + i2 * d2 + v] = src[i0 * s0 + i1 * s1 + i2 * s2 + v];
  }
}

// dst[idx[r], :] += src[r, :]  (bf16, rows of `vec_per_row` 16-byte vectors): the embedding
// weight gradient scattered straight into an existing dense gradient -- no [vocab, d] zero-fill,
// no sort, no second dense add.  One warp per token row, packed bf16x2 atomics (red.global).
__global__ void __launch_bounds__(256)
rows_scatter_add_bf16_kernel(const uint4* __restrict__ src, __nv_bfloat162* __restrict__ dst,
                             const long* __restrict__ idx, int n_rows, int vec_per_row,
                             long dst_rows) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const long t = idx[r];
    if (t < 0 || t >= dst_rows) continue;            // padding / ignore index
    __nv_bfloat162* drow = dst + t * vec_per_row * 4;
    const uint4* srow = src + static_cast<long>(r) * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) {
      const uint4 x = srow[v];
      const __nv_bfloat162* xs = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(drow + v * 4 + q, xs[q]);
    }
  }
}

}  // namespace

void launch_rows_scatter_add_bf16(const void* src, void* dst, const long* idx, int n_rows,
                                  int row_bytes, long dst_rows, cudaStream_t stream) {
  if (n_rows <= 0) return;
  int blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rows_scatter_add_bf16_kernel<<<blocks, 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<__nv_bfloat162*>(dst), idx, n_rows,
      row_bytes / 16, dst_rows);
}

void launch_rows_copy3(const void* const* src, void* const* dst, int n0, int n1, int n2,
                       int row_bytes, long s0, long s1, long s2, long d0, long d1, long d2,
                       cudaStream_t stream) {
  const long total = static_cast<long>(n0) * n1 * n2 * (row_bytes / 16);
  if (total <= 0) return;
  long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  Rows3 p;
  for (int i = 0; i < 3; ++i) {
    p.src[i] = reinterpret_cast<const uint4*>(src[i]);
    p.dst[i] = reinterpret_cast<uint4*>(dst[i]);
  }
  rows_copy3_kernel<<<dim3(static_cast<int>(blocks), 3), 256, 0, stream>>>(
      p, n0, n1, n2, row_bytes / 16, s0 / 16, s1 / 16, s2 / 16, d0 / 16, d1 / 16, d2 / 16);
}

void launch_rows_copy(const void* src, void* dst, int n0, int n1, int n2, int row_bytes,
                      long s0, long s1, long s2, long d0, long d1, long d2, cudaStream_t stream) {
  const long total = static_cast<long>(n0) * n1 * n2 * (row_bytes / 16);
  if (total <= 0) return;
  long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rows_copy_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), n0, n1, n2,
      row_bytes / 16, s0 / 16, s1 / 16, s2 / 16, d0 / 16, d1 / 16, d2 / 16);
}

}  // namespace tdp
