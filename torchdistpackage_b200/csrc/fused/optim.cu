// Flat-buffer optimizer / utility kernels (HBM-bound; 16-byte vectorised grid-stride loops).
//   adamw        : fused (un)scale + Adam/AdamW on an fp32 master shard + bf16 write-out (+ second
//                  copy into the symmetric all-gather slot) -- replaces optim.step() + master->bf16
//                  copy of the reference (ddp/zero_optim.py:257-277)
//   ema_update   : ema = ema*d + p*(1-d), flat or multi-tensor (dist/sharded_ema.py:21-31)
//   sumsq/scale  : gradient-norm clipping building blocks (pipeline_parallel/clip_grad_parallel.py)
//   cast_copy    : dtype-converting copy with scale (bucket pack / grad -> master grad)
#include "../common/ptx.cuh"
#include "../common/tdp_api.h"

namespace tdp {

namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t n_items, int per_thread = 4) {
  size_t blocks = (n_items + static_cast<size_t>(kThreads) * per_thread - 1) /
                  (static_cast<size_t>(kThreads) * per_thread);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return static_cast<int>(blocks);
}

template <bool kBf16>
__device__ __forceinline__ void load4(const void* base, size_t i4, float (&v)[4]) {
  if constexpr (kBf16) {
    const uint2 u = reinterpret_cast<const uint2*>(base)[i4];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    const float4 f = reinterpret_cast<const float4*>(base)[i4];
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
}
template <bool kBf16>
__device__ __forceinline__ void store4(void* base, size_t i4, const float (&v)[4]) {
  if constexpr (kBf16) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    reinterpret_cast<uint2*>(base)[i4] = u;
  } else {
    reinterpret_cast<float4*>(base)[i4] = make_float4(v[0], v[1], v[2], v[3]);
  }
}
template <bool kBf16>
__device__ __forceinline__ float load1(const void* base, size_t i) {
  if constexpr (kBf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  else return reinterpret_cast<const float*>(base)[i];
}
template <bool kBf16>
__device__ __forceinline__ void store1(void* base, size_t i, float v) {
  if constexpr (kBf16) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(base)[i] = v;
}

struct AdamConsts {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, gscale;
  int adamw;
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v,
                                          const AdamConsts& c) {
  if (!c.adamw) g += c.wd * p;          // classic L2
  else p *= (1.f - c.lr * c.wd);        // decoupled decay
  m = c.beta1 * m + (1.f - c.beta1) * g;
  v = c.beta2 * v + (1.f - c.beta2) * g * g;
  const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  return p - (c.lr / c.bc1) * (m / denom);
}

template <bool kParamBf16, bool kGradBf16, bool kHasMaster>
__global__ void __launch_bounds__(kThreads)
adamw_kernel(void* __restrict__ param, float* __restrict__ master, const void* __restrict__ grad,
             float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, size_t numel,
             AdamConsts c, const float* __restrict__ gscale_ptr, void* __restrict__ copy_out,
             int aligned, const float* __restrict__ hyper) {
  if (gscale_ptr) c.gscale *= *gscale_ptr;
  if (hyper) {
    // device-resident step / lr (CUDA-graph replay: the host values are frozen at capture time)
    const float step = hyper[0];
    c.lr = hyper[1];
    c.bc1 = 1.f - powf(c.beta1, step);
    c.bc2_sqrt = sqrtf(1.f - powf(c.beta2, step));
  }
  const size_t n4 = aligned ? numel / 4 : 0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += stride) {
    float p[4], g[4], m[4], v[4];
    if constexpr (kHasMaster) load4<false>(master, i, p);
    else load4<kParamBf16>(param, i, p);
    load4<kGradBf16>(grad, i, g);
    load4<false>(exp_avg, i, m);
    load4<false>(exp_avg_sq, i, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = adam_one(p[k], g[k] * c.gscale, m[k], v[k], c);
    store4<false>(exp_avg, i, m);
    store4<false>(exp_avg_sq, i, v);
    if constexpr (kHasMaster) store4<false>(master, i, p);
    store4<kParamBf16>(param, i, p);
    if (copy_out) store4<true>(copy_out, i, p);
  }
  // tail
  for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += stride) {
    float p = kHasMaster ? master[i] : load1<kParamBf16>(param, i);
    float g = load1<kGradBf16>(grad, i) * c.gscale;
    float m = exp_avg[i], v = exp_avg_sq[i];
    p = adam_one(p, g, m, v, c);
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    if (kHasMaster) master[i] = p;
    store1<kParamBf16>(param, i, p);
    if (copy_out) store1<true>(copy_out, i, p);
  }
}

template <bool kEmaBf16, bool kParamBf16>
__device__ __forceinline__ void ema_range(void* ema, const void* param, size_t numel, float decay,
                                          size_t tid, size_t stride) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(ema) | reinterpret_cast<uintptr_t>(param)) & 15) == 0;
  const size_t n4 = aligned ? numel / 4 : 0;
  for (size_t i = tid; i < n4; i += stride) {
    float e[4], p[4];
    load4<kEmaBf16>(ema, i, e);
    load4<kParamBf16>(param, i, p);
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = e[k] * decay + p[k] * (1.f - decay);
    store4<kEmaBf16>(ema, i, e);
  }
  for (size_t i = n4 * 4 + tid; i < numel; i += stride) {
    const float e = load1<kEmaBf16>(ema, i) * decay + load1<kParamBf16>(param, i) * (1.f - decay);
    store1<kEmaBf16>(ema, i, e);
  }
}

template <bool kParamBf16>
__global__ void __launch_bounds__(kThreads)
ema_kernel(float* ema, const void* param, size_t numel, float decay) {
  ema_range<false, kParamBf16>(ema, param, numel, decay,
                               static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x,
                               static_cast<size_t>(gridDim.x) * blockDim.x);
}

// one launch for a whole list of tensors: blockIdx.y = tensor index
__global__ void __launch_bounds__(kThreads)
ema_multi_kernel(void* const* ema_ptrs, const void* const* param_ptrs, const int64_t* numels,
                 const int* ema_dtypes, const int* param_dtypes, float decay) {
  const int t = blockIdx.y;
  const size_t n = static_cast<size_t>(numels[t]);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  void* e = ema_ptrs[t];
  const void* p = param_ptrs[t];
  const int ed = ema_dtypes[t], pd = param_dtypes[t];   // 0 = bf16, 1 = fp32
  if (ed == 1 && pd == 1) ema_range<false, false>(e, p, n, decay, tid, stride);
  else if (ed == 1 && pd == 0) ema_range<false, true>(e, p, n, decay, tid, stride);
  else if (ed == 0 && pd == 0) ema_range<true, true>(e, p, n, decay, tid, stride);
  else ema_range<true, false>(e, p, n, decay, tid, stride);
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const void* x, size_t numel, float* out) {
  float acc = 0.f;
  const size_t n4 = numel / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < n4; i += stride) {
    float v[4];
    load4<kBf16>(x, i, v);
    acc += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  for (size_t i = n4 * 4 + tid; i < numel; i += stride) {
    const float v = load1<kBf16>(x, i);
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < kThreads / 32 ? s[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads)
scale_kernel(void* x, size_t numel, float scale, const float* scale_ptr) {
  if (scale_ptr) scale *= *scale_ptr;
  const size_t n4 = numel / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < n4; i += stride) {
    float v[4];
    load4<kBf16>(x, i, v);
    v[0] *= scale; v[1] *= scale; v[2] *= scale; v[3] *= scale;
    store4<kBf16>(x, i, v);
  }
  for (size_t i = n4 * 4 + tid; i < numel; i += stride)
    store1<kBf16>(x, i, load1<kBf16>(x, i) * scale);
}

// multi-tensor variants (blockIdx.y = tensor index): one launch for a whole gradient list
// (clip_grad_norm_: reference parallel/pipeline_parallel/clip_grad_parallel.py:40-77 loops in Python)
template <bool kBf16>
__device__ __forceinline__ float sumsq_range(const void* x, size_t numel, size_t tid, size_t stride) {
  float acc = 0.f;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const size_t n4 = aligned ? numel / 4 : 0;
#pragma unroll 4
  for (size_t i = tid; i < n4; i += stride) {
    float v[4];
    load4<kBf16>(x, i, v);
    acc += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  for (size_t i = n4 * 4 + tid; i < numel; i += stride) {
    const float v = load1<kBf16>(x, i);
    acc += v * v;
  }
  return acc;
}

__global__ void __launch_bounds__(kThreads)
sumsq_multi_kernel(const void* const* ptrs, const int64_t* numels, const int* dtypes, float* out) {
  const int t = blockIdx.y;
  const size_t n = static_cast<size_t>(numels[t]);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  float acc = dtypes[t] == 0 ? sumsq_range<true>(ptrs[t], n, tid, stride)
                             : sumsq_range<false>(ptrs[t], n, tid, stride);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0.f) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(kThreads)
scale_multi_kernel(void* const* ptrs, const int64_t* numels, const int* dtypes, float scale,
                   const float* scale_ptr) {
  if (scale_ptr) scale *= *scale_ptr;
  const int t = blockIdx.y;
  const size_t n = static_cast<size_t>(numels[t]);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  void* x = ptrs[t];
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const size_t n4 = aligned ? n / 4 : 0;
  if (dtypes[t] == 0) {
    for (size_t i = tid; i < n4; i += stride) {
      float v[4];
      load4<true>(x, i, v);
      v[0] *= scale; v[1] *= scale; v[2] *= scale; v[3] *= scale;
      store4<true>(x, i, v);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) store1<true>(x, i, load1<true>(x, i) * scale);
  } else {
    for (size_t i = tid; i < n4; i += stride) {
      float v[4];
      load4<false>(x, i, v);
      v[0] *= scale; v[1] *= scale; v[2] *= scale; v[3] *= scale;
      store4<false>(x, i, v);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) store1<false>(x, i, load1<false>(x, i) * scale);
  }
}

template <bool kDstBf16, bool kSrcBf16>
__global__ void __launch_bounds__(kThreads)
cast_copy_kernel(void* dst, const void* src, size_t numel, float scale) {
  const size_t n4 = numel / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < n4; i += stride) {
    float v[4];
    load4<kSrcBf16>(src, i, v);
    v[0] *= scale; v[1] *= scale; v[2] *= scale; v[3] *= scale;
    store4<kDstBf16>(dst, i, v);
  }
  for (size_t i = n4 * 4 + tid; i < numel; i += stride)
    store1<kDstBf16>(dst, i, load1<kSrcBf16>(src, i) * scale);
}

}  // namespace

void launch_adamw(const AdamWLaunch& a, cudaStream_t stream) {
  if (a.numel == 0) return;
  AdamConsts c;
  c.lr = a.lr; c.beta1 = a.beta1; c.beta2 = a.beta2; c.eps = a.eps; c.wd = a.weight_decay;
  c.bc1 = a.bias_correction1; c.bc2_sqrt = sqrtf(a.bias_correction2); c.gscale = a.grad_scale;
  c.adamw = a.adamw_mode;
  const int grid = grid_for(a.numel / 4 + 1, 2);
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a.param) | reinterpret_cast<uintptr_t>(a.master) |
                         reinterpret_cast<uintptr_t>(a.grad) | reinterpret_cast<uintptr_t>(a.exp_avg) |
                         reinterpret_cast<uintptr_t>(a.exp_avg_sq) |
                         reinterpret_cast<uintptr_t>(a.param_copy_out);
  const int aligned = (bits & 15) == 0;
#define TDP_ADAM(PB, GB, HM)                                                              \
  adamw_kernel<PB, GB, HM><<<grid, kThreads, 0, stream>>>(a.param, a.master, a.grad,      \
      a.exp_avg, a.exp_avg_sq, a.numel, c, a.grad_scale_ptr, a.param_copy_out, aligned, a.hyper)
  const bool hm = a.master != nullptr;
  if (a.param_bf16 && a.grad_bf16) { if (hm) TDP_ADAM(true, true, true); else TDP_ADAM(true, true, false); }
  else if (a.param_bf16 && !a.grad_bf16) { if (hm) TDP_ADAM(true, false, true); else TDP_ADAM(true, false, false); }
  else if (!a.param_bf16 && a.grad_bf16) { if (hm) TDP_ADAM(false, true, true); else TDP_ADAM(false, true, false); }
  else { if (hm) TDP_ADAM(false, false, true); else TDP_ADAM(false, false, false); }
#undef TDP_ADAM
}

void launch_ema_update(float* ema, const void* param, int param_bf16, size_t numel, float decay,
                       cudaStream_t stream) {
  if (numel == 0) return;
  const int grid = grid_for(numel / 4 + 1, 2);
  if (param_bf16) ema_kernel<true><<<grid, kThreads, 0, stream>>>(ema, param, numel, decay);
  else ema_kernel<false><<<grid, kThreads, 0, stream>>>(ema, param, numel, decay);
}

void launch_ema_update_multi(void* const* ema_ptrs, const void* const* param_ptrs,
                             const int64_t* numels, const int* ema_dtypes, const int* param_dtypes,
                             int n_tensors, float decay, cudaStream_t stream) {
  if (n_tensors <= 0) return;
  dim3 grid(16, n_tensors);
  ema_multi_kernel<<<grid, kThreads, 0, stream>>>(ema_ptrs, param_ptrs, numels, ema_dtypes,
                                                  param_dtypes, decay);
}

void launch_sumsq(const void* x, int dtype, size_t numel, float* out, cudaStream_t stream) {
  if (numel == 0) return;
  const int grid = grid_for(numel / 4 + 1, 4);
  if (dtype == 0) sumsq_kernel<true><<<grid, kThreads, 0, stream>>>(x, numel, out);
  else sumsq_kernel<false><<<grid, kThreads, 0, stream>>>(x, numel, out);
}

void launch_scale_(void* x, int dtype, size_t numel, float scale, const float* scale_ptr,
                   cudaStream_t stream) {
  if (numel == 0) return;
  const int grid = grid_for(numel / 4 + 1, 2);
  if (dtype == 0) scale_kernel<true><<<grid, kThreads, 0, stream>>>(x, numel, scale, scale_ptr);
  else scale_kernel<false><<<grid, kThreads, 0, stream>>>(x, numel, scale, scale_ptr);
}

void launch_sumsq_multi(const void* const* ptrs, const int64_t* numels, const int* dtypes,
                        int n_tensors, float* out, cudaStream_t stream) {
  if (n_tensors <= 0) return;
  dim3 grid(148, n_tensors);         // one CTA column per SM: enough loads in flight for the big tensors
  sumsq_multi_kernel<<<grid, kThreads, 0, stream>>>(ptrs, numels, dtypes, out);
}

void launch_scale_multi(void* const* ptrs, const int64_t* numels, const int* dtypes, int n_tensors,
                        float scale, const float* scale_ptr, cudaStream_t stream) {
  if (n_tensors <= 0) return;
  dim3 grid(148, n_tensors);
  scale_multi_kernel<<<grid, kThreads, 0, stream>>>(ptrs, numels, dtypes, scale, scale_ptr);
}

void launch_cast_copy(void* dst, int dst_dtype, const void* src, int src_dtype, size_t numel,
                      float scale, cudaStream_t stream) {
  if (numel == 0) return;
  const int grid = grid_for(numel / 4 + 1, 2);
  if (dst_dtype == 0 && src_dtype == 0) cast_copy_kernel<true, true><<<grid, kThreads, 0, stream>>>(dst, src, numel, scale);
  else if (dst_dtype == 0) cast_copy_kernel<true, false><<<grid, kThreads, 0, stream>>>(dst, src, numel, scale);
  else if (src_dtype == 0) cast_copy_kernel<false, true><<<grid, kThreads, 0, stream>>>(dst, src, numel, scale);
  else cast_copy_kernel<false, false><<<grid, kThreads, 0, stream>>>(dst, src, numel, scale);
}

}  // namespace tdp
