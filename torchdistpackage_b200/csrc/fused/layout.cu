// Layout kernels around attention: strided 3-D row copies with 16-byte vectors so that the
// [B,H,T,D] <-> [B,T,H,D] (and packed qkv gradient) permutes run at HBM speed instead of through
// generic strided elementwise / cat kernels.
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

}  // namespace

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
