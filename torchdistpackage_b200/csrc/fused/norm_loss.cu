// Row-wise fused kernels (HBM-bound): LayerNorm fwd/bwd with fused residual add, column sums
// (bias gradients) and fused softmax-cross-entropy forward+backward.
//   LayerNorm replaces nn.LayerNorm + the separate residual add of the reference Block
//   (parallel/tensor_parallel/transformer.py:19-35).
#include "../common/ptx.cuh"
#include "../common/tdp_api.h"

namespace tdp {

namespace {

constexpr int kLnThreads = 256;
constexpr int kMaxVpt = 4;  // vectors (8 bf16) per thread -> cols <= 256*4*8 = 8192

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int w = threadIdx.x / 32, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  float r = (l < kThreads / 32) ? smem[l] : 0.f;
  r = warp_sum(r);
  return r;
}
template <int kThreads>
__device__ __forceinline__ float block_max(float v, float* smem) {
  v = warp_max(v);
  const int w = threadIdx.x / 32, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  float r = (l < kThreads / 32) ? smem[l] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// y = LN(x (+ residual)) * gamma + beta ; optionally writes the pre-norm sum (resid_out)
__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                     const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ resid_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int cols,
                     float eps) {
  __shared__ float red[kLnThreads / 32];
  const int nvec = cols / 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t base = static_cast<size_t>(row) * cols;
    float v[kMaxVpt][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVpt; ++i) {
      const int vi = threadIdx.x + i * kLnThreads;
      if (vi < nvec) {
        unpack8(*reinterpret_cast<const uint4*>(x + base + vi * 8), v[i]);
        if (residual) {
          float r[8];
          unpack8(*reinterpret_cast<const uint4*>(residual + base + vi * 8), r);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[i][k] += r[k];
        }
        if (resid_out) {
          // round through bf16 so that fwd statistics match what backward will re-read
          const uint4 packed = pack8f(v[i]);
          *reinterpret_cast<uint4*>(resid_out + base + vi * 8) = packed;
          unpack8(packed, v[i]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += v[i][k];
      }
    }
    const float mean = block_sum<kLnThreads>(sum, red) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVpt; ++i) {
      const int vi = threadIdx.x + i * kLnThreads;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; sq += d * d; }
      }
    }
    const float var = block_sum<kLnThreads>(sq, red) / cols;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < kMaxVpt; ++i) {
      const int vi = threadIdx.x + i * kLnThreads;
      if (vi < nvec) {
        float g[8], b[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), g);
        if (beta) unpack8(*reinterpret_cast<const uint4*>(beta + vi * 8), b);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          o[k] = (v[i][k] - mean) * rstd * g[k] + (beta ? b[k] : 0.f);
        *reinterpret_cast<uint4*>(y + base + vi * 8) = pack8f(o);
      }
    }
  }
}

// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)) (+ dresid) ;
// per-block partial dgamma / dbeta -> [gridDim.x, cols] fp32
__global__ void __launch_bounds__(kLnThreads)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx,
                     const __nv_bfloat16* __restrict__ dresid, float* __restrict__ dgamma_partial,
                     float* __restrict__ dbeta_partial, int rows, int cols) {
  __shared__ float red[kLnThreads / 32];
  const int nvec = cols / 8;
  float dg[kMaxVpt][8], db[kMaxVpt][8], g[kMaxVpt][8];
#pragma unroll
  for (int i = 0; i < kMaxVpt; ++i) {
    const int vi = threadIdx.x + i * kLnThreads;
#pragma unroll
    for (int k = 0; k < 8; ++k) { dg[i][k] = 0.f; db[i][k] = 0.f; g[i][k] = 0.f; }
    if (vi < nvec) unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), g[i]);
  }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t base = static_cast<size_t>(row) * cols;
    const float mu = mean[row], rs = rstd[row];
    float xh[kMaxVpt][8], dyv[kMaxVpt][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVpt; ++i) {
      const int vi = threadIdx.x + i * kLnThreads;
      if (vi < nvec) {
        unpack8(*reinterpret_cast<const uint4*>(x + base + vi * 8), xh[i]);
        unpack8(*reinterpret_cast<const uint4*>(dy + base + vi * 8), dyv[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          xh[i][k] = (xh[i][k] - mu) * rs;
          dg[i][k] += dyv[i][k] * xh[i][k];
          db[i][k] += dyv[i][k];
          const float t = dyv[i][k] * g[i][k];
          s1 += t;
          s2 += t * xh[i][k];
        }
      }
    }
    s1 = block_sum<kLnThreads>(s1, red) / cols;
    s2 = block_sum<kLnThreads>(s2, red) / cols;
#pragma unroll
    for (int i = 0; i < kMaxVpt; ++i) {
      const int vi = threadIdx.x + i * kLnThreads;
      if (vi < nvec) {
        float o[8], r[8];
        if (dresid) unpack8(*reinterpret_cast<const uint4*>(dresid + base + vi * 8), r);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o[k] = rs * (dyv[i][k] * g[i][k] - s1 - xh[i][k] * s2);
          if (dresid) o[k] += r[k];
        }
        *reinterpret_cast<uint4*>(dx + base + vi * 8) = pack8f(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxVpt; ++i) {
    const int vi = threadIdx.x + i * kLnThreads;
    if (vi < nvec) {
      float* pg = dgamma_partial + static_cast<size_t>(blockIdx.x) * cols + vi * 8;
      float* pb = dbeta_partial + static_cast<size_t>(blockIdx.x) * cols + vi * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) { pg[k] = dg[i][k]; pb[k] = db[i][k]; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-row LayerNorm (cols = 256 * kVpl, kVpl <= 4): no block barriers, shuffle reductions only,
// several rows in flight per SM -> HBM-bound instead of latency-bound.
// ---------------------------------------------------------------------------------------------
constexpr int kLnWarps = 4;

template <int kVpl>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_warp_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                          const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                          __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ resid_out,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                          float eps) {
  constexpr int cols = kVpl * 256;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kLnWarps + threadIdx.x / 32;
  const int n_warps = gridDim.x * kLnWarps;
  float g[kVpl][8], b[kVpl][8];
#pragma unroll
  for (int i = 0; i < kVpl; ++i) {
    unpack8(*reinterpret_cast<const uint4*>(gamma + (lane + 32 * i) * 8), g[i]);
    if (beta) unpack8(*reinterpret_cast<const uint4*>(beta + (lane + 32 * i) * 8), b[i]);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = 0.f;
    }
  }
  for (int row = warp; row < rows; row += n_warps) {
    const size_t base = static_cast<size_t>(row) * cols;
    float v[kVpl][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVpl; ++i) {
      const int off = (lane + 32 * i) * 8;
      unpack8(*reinterpret_cast<const uint4*>(x + base + off), v[i]);
      if (residual) {
        float r[8];
        unpack8(*reinterpret_cast<const uint4*>(residual + base + off), r);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] += r[k];
      }
      if (resid_out) {
        const uint4 packed = pack8f(v[i]);
        *reinterpret_cast<uint4*>(resid_out + base + off) = packed;
        unpack8(packed, v[i]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[i][k];
    }
    const float mean = warp_sum(sum) * (1.f / cols);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kVpl; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; sq += d * d; }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / cols) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < kVpl; ++i) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[i][k] + b[i][k];
      *reinterpret_cast<uint4*>(y + base + (lane + 32 * i) * 8) = pack8f(o);
    }
  }
}

template <int kVpl>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_warp_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx,
                          const __nv_bfloat16* __restrict__ dresid, float* __restrict__ dgamma_partial,
                          float* __restrict__ dbeta_partial, int rows) {
  constexpr int cols = kVpl * 256;
  __shared__ float sred[kLnWarps][cols];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x / 32;
  const int warp = blockIdx.x * kLnWarps + wib;
  const int n_warps = gridDim.x * kLnWarps;
  float g[kVpl][8], dg[kVpl][8], db[kVpl][8];
#pragma unroll
  for (int i = 0; i < kVpl; ++i) {
    unpack8(*reinterpret_cast<const uint4*>(gamma + (lane + 32 * i) * 8), g[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) { dg[i][k] = 0.f; db[i][k] = 0.f; }
  }
  for (int row = warp; row < rows; row += n_warps) {
    const size_t base = static_cast<size_t>(row) * cols;
    const float mu = mean[row], rs = rstd[row];
    float xh[kVpl][8], dyv[kVpl][8];
    uint4 xr[kVpl], dr[kVpl], rr[kVpl];
    float s1 = 0.f, s2 = 0.f;
    // every load of the row is issued before the first use (the skip gradient included: loading
    // it after the two reductions cost a third exposed memory round trip per row)
#pragma unroll
    for (int i = 0; i < kVpl; ++i) {
      const int off = (lane + 32 * i) * 8;
      xr[i] = *reinterpret_cast<const uint4*>(x + base + off);
      dr[i] = *reinterpret_cast<const uint4*>(dy + base + off);
      if (dresid) rr[i] = *reinterpret_cast<const uint4*>(dresid + base + off);
    }
#pragma unroll
    for (int i = 0; i < kVpl; ++i) {
      unpack8(xr[i], xh[i]);
      unpack8(dr[i], dyv[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[i][k] = (xh[i][k] - mu) * rs;
        dg[i][k] += dyv[i][k] * xh[i][k];
        db[i][k] += dyv[i][k];
        const float t = dyv[i][k] * g[i][k];
        s1 += t;
        s2 += t * xh[i][k];
      }
    }
    s1 = warp_sum(s1) * (1.f / cols);
    s2 = warp_sum(s2) * (1.f / cols);
#pragma unroll
    for (int i = 0; i < kVpl; ++i) {
      const int off = (lane + 32 * i) * 8;
      float o[8], r[8];
      if (dresid) unpack8(rr[i], r);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o[k] = rs * (dyv[i][k] * g[i][k] - s1 - xh[i][k] * s2);
        if (dresid) o[k] += r[k];
      }
      *reinterpret_cast<uint4*>(dx + base + off) = pack8f(o);
    }
  }
  // block-level reduction of the per-warp dgamma / dbeta partials (two passes through smem)
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kVpl; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sred[wib][(lane + 32 * i) * 8 + k] = pass == 0 ? dg[i][k] : db[i][k];
    __syncthreads();
    float* outp = (pass == 0 ? dgamma_partial : dbeta_partial) + static_cast<size_t>(blockIdx.x) * cols;
    for (int c = threadIdx.x; c < cols; c += kLnWarps * 32) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) acc += sred[w][c];
      outp[c] = acc;
    }
  }
}

// column sums with one pass and enough loads in flight: block = 8 warps x (32 lanes x 8 columns);
// warp w walks rows r0+w, r0+w+8, ...; partials meet in smem, then one fp32 atomicAdd per column.
constexpr int kCsRows = 64;
__global__ void __launch_bounds__(256)
colsum_atomic_kernel(const __nv_bfloat16* __restrict__ x, int rows, int cols, int ld,
                     float* __restrict__ scratch /* persistent fp32 [cols], all zero on entry */,
                     unsigned int* __restrict__ tickets /* one per column group, zero on entry */,
                     void* __restrict__ out, int out_bf16) {
  __shared__ float sred[8][256];
  __shared__ unsigned int s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x / 32;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * kCsRows;
  const int r1 = min(rows, r0 + kCsRows);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < cols) {
#pragma unroll 4
    for (int r = r0 + w; r < r1; r += 8) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld + col), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sred[w][lane * 8 + k] = acc[k];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += sred[ww][threadIdx.x];
    atomicAdd(scratch + c, t);
  }
  // the last block of this column group converts the sums and leaves scratch / ticket zeroed for
  // the next call: one launch per bias gradient, no memset, no second kernel
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (c < cols) {
      const float t = atomicExch(scratch + c, 0.f);
      if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[c] = __float2bfloat16_rn(t);
      else reinterpret_cast<float*>(out)[c] = t;
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
  }
}

// out[c] = sum_p partial[p][c]: block = 8 warps x 32 columns; warp w sums partials w, w+8, ...
// (coalesced 128-byte rows), then the 8 warp sums meet in shared memory.
constexpr int kPrWarps = 32;     // 1024 threads: the kernel is a handful of CTAs deep, so the
                                 // loads in flight per CTA are what bounds it (was 8 warps: 21 us)
__global__ void __launch_bounds__(32 * kPrWarps)
colsum_partial_reduce_kernel(const float* __restrict__ partial, int n_partial, int cols,
                             void* __restrict__ out, int out_bf16, size_t partial_stride_y,
                             void* __restrict__ out_y1) {
  // blockIdx.y selects an independent problem (LayerNorm: 0 = dgamma, 1 = dbeta)
  __shared__ float sred[kPrWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + lane;
  partial += blockIdx.y * partial_stride_y;
  if (blockIdx.y == 1) out = out_y1;
  float acc = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int p = w; p < n_partial; p += kPrWarps) acc += partial[static_cast<size_t>(p) * cols + c];
  }
  sred[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kPrWarps; ++i) t += sred[i][lane];
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[c] = __float2bfloat16_rn(t);
    else reinterpret_cast<float*>(out)[c] = t;
  }
}

// partial column sums: block b handles rows [b*rows_per_block, ...), thread = 8 columns
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int rows, int cols, int ld,
                      int rows_per_block, float* __restrict__ partial) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi * 8 >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld + vi * 8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
  float* p = partial + static_cast<size_t>(blockIdx.y) * cols + vi * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) p[k] = acc[k];
}

// One block per row.  Pass 1: online max / sum-exp.  Pass 2: dlogits = (softmax - onehot) * gs
// written in place.  loss[row] = lse - logit[target]  (0 for ignored rows).
constexpr int kCeThreads = 512;
__global__ void __launch_bounds__(kCeThreads)
cross_entropy_kernel(__nv_bfloat16* __restrict__ logits, int rows, int vocab, int ld,
                     const int64_t* __restrict__ target, float* __restrict__ loss, float grad_scale,
                     int ignore_index) {
  __shared__ float red[kCeThreads / 32];
  const int row = blockIdx.x;
  __nv_bfloat16* lrow = logits + static_cast<size_t>(row) * ld;
  const int64_t tgt = target[row];
  const int nvec = vocab / 8;
  if (tgt == ignore_index) {
    for (int vi = threadIdx.x; vi < nvec; vi += kCeThreads)
      *reinterpret_cast<uint4*>(lrow + vi * 8) = make_uint4(0, 0, 0, 0);
    for (int c = nvec * 8 + threadIdx.x; c < vocab; c += kCeThreads) lrow[c] = __float2bfloat16_rn(0.f);
    if (threadIdx.x == 0) loss[row] = 0.f;
    return;
  }
  float m = -INFINITY, s = 0.f;
  for (int vi = threadIdx.x; vi < nvec; vi += kCeThreads) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(lrow + vi * 8), v);
    float lm = v[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) lm = fmaxf(lm, v[k]);
    const float nm = fmaxf(m, lm);
    float add = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) add += __expf(v[k] - nm);
    s = s * __expf(m - nm) + add;
    m = nm;
  }
  for (int c = nvec * 8 + threadIdx.x; c < vocab; c += kCeThreads) {
    const float v = __bfloat162float(lrow[c]);
    const float nm = fmaxf(m, v);
    s = s * __expf(m - nm) + __expf(v - nm);
    m = nm;
  }
  const float gm = block_max<kCeThreads>(m, red);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_sum<kCeThreads>(s, red);
  const float lse = gm + __logf(gs);
  const float tv = __bfloat162float(lrow[tgt]);
  __syncthreads();   // everyone has read the target logit before it is overwritten
  if (threadIdx.x == 0) loss[row] = lse - tv;
  const float inv = 1.f / gs;
  for (int vi = threadIdx.x; vi < nvec; vi += kCeThreads) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(lrow + vi * 8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float p = __expf(v[k] - gm) * inv;
      if (vi * 8 + k == tgt) p -= 1.f;
      v[k] = p * grad_scale;
    }
    *reinterpret_cast<uint4*>(lrow + vi * 8) = pack8f(v);
  }
  for (int c = nvec * 8 + threadIdx.x; c < vocab; c += kCeThreads) {
    float p = __expf(__bfloat162float(lrow[c]) - gm) * inv;
    if (c == tgt) p -= 1.f;
    lrow[c] = __float2bfloat16_rn(p * grad_scale);
  }
}

}  // namespace

void launch_layernorm_fwd(const void* x, const void* residual, const void* gamma, const void* beta,
                          void* y, void* resid_out, float* mean, float* rstd, int rows, int cols,
                          float eps, cudaStream_t stream) {
  if (rows <= 0) return;
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto rb = reinterpret_cast<const __nv_bfloat16*>(residual);
  auto gb = reinterpret_cast<const __nv_bfloat16*>(gamma);
  auto bb = reinterpret_cast<const __nv_bfloat16*>(beta);
  auto yb = reinterpret_cast<__nv_bfloat16*>(y);
  auto ro = reinterpret_cast<__nv_bfloat16*>(resid_out);
  if (cols % 256 == 0 && cols <= 1024) {
    int grid = (rows + kLnWarps - 1) / kLnWarps;
    if (grid > 148 * 16) grid = 148 * 16;
    switch (cols / 256) {
      case 1: layernorm_fwd_warp_kernel<1><<<grid, kLnWarps * 32, 0, stream>>>(xb, rb, gb, bb, yb, ro, mean, rstd, rows, eps); break;
      case 2: layernorm_fwd_warp_kernel<2><<<grid, kLnWarps * 32, 0, stream>>>(xb, rb, gb, bb, yb, ro, mean, rstd, rows, eps); break;
      case 3: layernorm_fwd_warp_kernel<3><<<grid, kLnWarps * 32, 0, stream>>>(xb, rb, gb, bb, yb, ro, mean, rstd, rows, eps); break;
      default: layernorm_fwd_warp_kernel<4><<<grid, kLnWarps * 32, 0, stream>>>(xb, rb, gb, bb, yb, ro, mean, rstd, rows, eps); break;
    }
    return;
  }
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  layernorm_fwd_kernel<<<grid, kLnThreads, 0, stream>>>(xb, rb, gb, bb, yb, ro, mean, rstd, rows, cols, eps);
}

void launch_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                          const float* rstd, void* dx, const void* dresid, float* dgamma_partial,
                          float* dbeta_partial, int rows, int cols, int n_partial,
                          cudaStream_t stream) {
  if (rows <= 0) return;
  auto dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto gb = reinterpret_cast<const __nv_bfloat16*>(gamma);
  auto dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  auto drb = reinterpret_cast<const __nv_bfloat16*>(dresid);
  if (cols % 256 == 0 && cols <= 1024) {
    switch (cols / 256) {
      case 1: layernorm_bwd_warp_kernel<1><<<n_partial, kLnWarps * 32, 0, stream>>>(dyb, xb, gb, mean, rstd, dxb, drb, dgamma_partial, dbeta_partial, rows); break;
      case 2: layernorm_bwd_warp_kernel<2><<<n_partial, kLnWarps * 32, 0, stream>>>(dyb, xb, gb, mean, rstd, dxb, drb, dgamma_partial, dbeta_partial, rows); break;
      case 3: layernorm_bwd_warp_kernel<3><<<n_partial, kLnWarps * 32, 0, stream>>>(dyb, xb, gb, mean, rstd, dxb, drb, dgamma_partial, dbeta_partial, rows); break;
      default: layernorm_bwd_warp_kernel<4><<<n_partial, kLnWarps * 32, 0, stream>>>(dyb, xb, gb, mean, rstd, dxb, drb, dgamma_partial, dbeta_partial, rows); break;
    }
    return;
  }
  layernorm_bwd_kernel<<<n_partial, kLnThreads, 0, stream>>>(dyb, xb, gb, mean, rstd, dxb, drb,
                                                            dgamma_partial, dbeta_partial, rows, cols);
}

void launch_colsum_partial_reduce(const float* partial, int n_partial, int cols, void* out,
                                  int out_bf16, cudaStream_t stream) {
  colsum_partial_reduce_kernel<<<(cols + 31) / 32, 32 * kPrWarps, 0, stream>>>(partial, n_partial, cols, out,
                                                                    out_bf16, 0, nullptr);
}

void launch_colsum_partial_reduce2(const float* partial, int n_partial, int cols, void* out0,
                                   void* out1, int out_bf16, cudaStream_t stream) {
  // partial = [2][n_partial][cols]; one launch reduces both halves
  dim3 grid((cols + 31) / 32, 2);
  colsum_partial_reduce_kernel<<<grid, 32 * kPrWarps, 0, stream>>>(
      partial, n_partial, cols, out0, out_bf16, static_cast<size_t>(n_partial) * cols, out1);
}

void launch_colsum(const void* x, int rows, int cols, int ld, float* /*scratch (unused)*/,
                   void* out, int out_bf16, cudaStream_t stream) {
  // persistent self-cleaning accumulators (per device): zero before and after every call
  constexpr int kMaxCols = 1 << 18;
  static float* acc[16] = {nullptr};
  static unsigned int* tick[16] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (acc[dev] == nullptr) {
    cudaMalloc(&acc[dev], sizeof(float) * kMaxCols);
    cudaMemset(acc[dev], 0, sizeof(float) * kMaxCols);
    cudaMalloc(&tick[dev], sizeof(unsigned int) * (kMaxCols / 256));
    cudaMemset(tick[dev], 0, sizeof(unsigned int) * (kMaxCols / 256));
  }
  for (int c0 = 0; c0 < cols; c0 += kMaxCols) {
    const int nc = cols - c0 < kMaxCols ? cols - c0 : kMaxCols;
    dim3 grid((nc + 255) / 256, (rows + kCsRows - 1) / kCsRows);
    colsum_atomic_kernel<<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x) + c0, rows, nc, ld, acc[dev], tick[dev],
        static_cast<char*>(out) + static_cast<size_t>(c0) * (out_bf16 ? 2 : 4), out_bf16);
  }
}

void launch_cross_entropy_fwd_bwd(void* logits, int rows, int vocab, int ld, const int64_t* target,
                                  float* loss, float grad_scale, int ignore_index,
                                  cudaStream_t stream) {
  if (rows <= 0) return;
  cross_entropy_kernel<<<rows, kCeThreads, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(logits),
                                                        rows, vocab, ld, target, loss, grad_scale,
                                                        ignore_index);
}

}  // namespace tdp
