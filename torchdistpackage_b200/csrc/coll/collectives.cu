// NVSwitch collectives over symmetric memory, written for 8xB200 (NVLink 5, NVLS multicast).
//
//   * all_reduce      two-shot: every rank reduces 1/N of the buffer -- in the switch with
//                     multimem.ld_reduce (NVLS) or by reading its slice from every peer (P2P) --
//                     and broadcasts the result with multimem.st / peer stores.
//   * reduce_scatter  first half of the above, result written to a local (bf16|fp32) tensor with an
//                     optional fp32 accumulate (ZeRO master-grad epilogue).
//   * all_gather      multimem.st (one store fans out in the switch) or peer stores of my slice.
//   * rs_reduce       second stage of the fused GEMM->reduce-scatter / GEMM->all-reduce: waits for
//                     the per-source tile counters bumped by the GEMM epilogues, sums the partials.
//   * a2a rows        MoE dispatch / combine: row-granular peer stores / loads.
//
// Cross-GPU synchronisation uses flag words in the symmetric signal pad: release/acquire CAS at
// system scope, one slot per (block, peer) so blocks synchronise independently (no grid sync).
//
// These replace the reference's NCCL calls: dist.all_reduce on DDP buckets (ddp/naive_ddp.py:104-127),
// ZeRO's all_reduce + per-tensor broadcast (ddp/zero_optim.py:73-95,282-287), TP/SP collectives
// (parallel/tensor_parallel/tp_utils.py:44,67,84).
#include <stdlib.h>

#include "../common/ptx.cuh"
#include "../common/tdp_api.h"

namespace tdp {

namespace {

// CTA shape of the collective kernels.  Measured on 2 x B200 inside the overlapped GPT-2 step
// (profiles/r2/ab_n2.txt): 512-thread CTAs (32 of them) beat 128-thread CTAs that co-reside with
// the persistent GEMM CTAs on every SM (18.94 vs 19.42 ms/step fused, 19.23 vs 19.46 plain) --
// fewer, fatter CTAs disturb fewer GEMM CTAs.  TDP_COLL_THREADS = 128 | 256 | 512 overrides.
constexpr int kCollThreadsMax = 512;        // __launch_bounds__ of the collective kernels
constexpr int kMaxCollBlocks = 128;         // barrier slots reserved per signal pad region
// CTA shape actually launched: TDP_COLL_THREADS (128 | 256 | 512), default 512
inline int coll_threads() {
  static const int v = [] {
    const char* e = getenv("TDP_COLL_THREADS");
    const int t = e ? atoi(e) : 512;
    return (t == 128 || t == 256) ? t : 512;
  }();
  return v;
}
#define kCollThreads coll_threads()
constexpr int kBarrierSlotWords = kMaxCollBlocks * kApiMaxPeers;   // words per barrier "slot"

struct PeerPtrs {
  void* buf[kApiMaxPeers];
  uint32_t* signal[kApiMaxPeers];
};

__device__ __forceinline__ void put_signal(uint32_t* addr) {
  uint32_t old;
  SpinWatchdog wd;
  for (;;) {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    if (old == 0u) break;
    wd.tick("barrier put_signal", threadIdx.x, 0);
  }
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {
  uint32_t old;
  SpinWatchdog wd;
  for (;;) {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (old == 1u) break;
    wd.tick("barrier wait_signal", threadIdx.x, 1);
  }
}

// Block-level barrier with the same-index block of every peer.  Slot layout (uint32 words):
//   pad[slot_base + blockIdx.x * world + src_rank]
__device__ __forceinline__ void block_barrier(const PeerPtrs& pp, int rank, int world,
                                              int slot_base) {
  // every thread drains its own peer / multimem stores to system scope before the block's
  // signalling threads publish: the release of the signalling thread alone is not relied upon
  // to cover the other threads' in-flight multicast stores (8-GPU fan-out is slow enough for a
  // peer to overwrite -- e.g. zero -- a slot and then receive a late store on top of it).
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    put_signal(pp.signal[peer] + slot_base + blockIdx.x * world + rank);
    wait_signal(pp.signal[rank] + slot_base + blockIdx.x * world + peer);
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 scale_bf16x8(uint4 v, float s) {
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z),
         d = unpack_bf16x2(v.w);
  v.x = pack_bf16x2(a.x * s, a.y * s);
  v.y = pack_bf16x2(b.x * s, b.y * s);
  v.z = pack_bf16x2(c.x * s, c.y * s);
  v.w = pack_bf16x2(d.x * s, d.y * s);
  return v;
}
__device__ __forceinline__ void acc_bf16x8(float (&acc)[8], uint4 v) {
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z),
         d = unpack_bf16x2(v.w);
  acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
  acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&acc)[8], float s) {
  uint4 v;
  v.x = pack_bf16x2(acc[0] * s, acc[1] * s);
  v.y = pack_bf16x2(acc[2] * s, acc[3] * s);
  v.z = pack_bf16x2(acc[4] * s, acc[5] * s);
  v.w = pack_bf16x2(acc[6] * s, acc[7] * s);
  return v;
}

// reduce 16 bytes at byte offset `off` across all ranks. kMc: in-switch, else peer loads (fixed
// rank order starting at rank 0 -> bitwise identical result on every rank).
template <bool kMc, bool kFp32>
__device__ __forceinline__ uint4 reduce16(const PeerPtrs& pp, const char* mc, int world,
                                          size_t off, float scale) {
  if constexpr (kMc) {
    if constexpr (kFp32) {
      float4 f = multimem_ld_reduce_f32x4(mc + off);
      f.x *= scale; f.y *= scale; f.z *= scale; f.w *= scale;
      return *reinterpret_cast<uint4*>(&f);
    } else {
      uint4 v = multimem_ld_reduce_bf16x8(mc + off);
      return scale == 1.f ? v : scale_bf16x8(v, scale);
    }
  } else {
    if constexpr (kFp32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kApiMaxPeers; ++p) {
        if (p < world) {
          uint4 u = ld_nc_v4(reinterpret_cast<const char*>(pp.buf[p]) + off);
          float4 f = *reinterpret_cast<float4*>(&u);
          acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
        }
      }
      acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
      return *reinterpret_cast<uint4*>(&acc);
    } else {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      uint4 vals[kApiMaxPeers];
#pragma unroll
      for (int p = 0; p < kApiMaxPeers; ++p)
        if (p < world) vals[p] = ld_nc_v4(reinterpret_cast<const char*>(pp.buf[p]) + off);
#pragma unroll
      for (int p = 0; p < kApiMaxPeers; ++p)
        if (p < world) acc_bf16x8(acc, vals[p]);
      return pack8(acc, scale);
    }
  }
}

// ------------------------------------------------------------------------------------------
// all-reduce (two-shot, in place)
// ------------------------------------------------------------------------------------------
template <bool kMc, bool kFp32>
__global__ void __launch_bounds__(kCollThreadsMax)
all_reduce_two_shot_kernel(const __grid_constant__ PeerPtrs pp, char* mc, int rank, int world, size_t offset,
                           size_t n_vec /*16-byte vectors*/, float scale, int slot_base) {
  block_barrier(pp, rank, world, slot_base);
  // my part: vectors [lo, hi)
  const size_t per = (n_vec + world - 1) / world;
  const size_t lo = min(per * rank, n_vec), hi = min(lo + per, n_vec);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  // kUnroll independent 16-byte reductions in flight per thread: NVLink round trips are ~2-4 us,
  // bandwidth = bytes in flight / latency.
  constexpr int kUnroll = kMc ? 8 : 4;
  for (size_t i = lo + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
       i += stride * kUnroll) {
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx < hi) v[u] = reduce16<kMc, kFp32>(pp, mc, world, offset + idx * 16, scale);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx < hi) {
        const size_t off = offset + idx * 16;
        if constexpr (kMc) {
          multimem_st_v4(mc + off, v[u]);
        } else {
#pragma unroll
          for (int p = 0; p < kApiMaxPeers; ++p)
            if (p < world) st_na_v4(reinterpret_cast<char*>(pp.buf[p]) + off, v[u]);
        }
      }
    }
  }
  block_barrier(pp, rank, world, slot_base + kBarrierSlotWords);
}

// ------------------------------------------------------------------------------------------
// all-reduce, one-shot (latency path for small messages): every rank reduces the WHOLE buffer
// itself -- one multimem.ld_reduce per 16 bytes (or a load from every peer) -- and keeps the
// result locally; no second shot, no cross-GPU stores.  The result goes to a private copy first
// and is written back after the closing barrier, so no rank overwrites data a peer still reads.
// Cost: N x the switch reads of the two-shot algorithm, one NVLink round trip less -- wins up to
// a few hundred KiB (reference: the many small dist.all_reduce calls of naive_ddp.py:158-160,171).
// ------------------------------------------------------------------------------------------
template <bool kMc, bool kFp32>
__global__ void __launch_bounds__(kCollThreadsMax)
all_reduce_one_shot_kernel(const __grid_constant__ PeerPtrs pp, const char* mc, int rank, int world,
                           size_t offset, size_t n_vec, float scale, uint4* scratch, int slot_base) {
  block_barrier(pp, rank, world, slot_base);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride)
    scratch[i] = reduce16<kMc, kFp32>(pp, mc, world, offset + i * 16, scale);
  block_barrier(pp, rank, world, slot_base + kBarrierSlotWords);     // every peer has read my data
  char* mine = reinterpret_cast<char*>(pp.buf[rank]) + offset;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride)
    reinterpret_cast<uint4*>(mine)[i] = scratch[i];
}

// ------------------------------------------------------------------------------------------
// reduce-scatter -> AdamW on my shard -> all-gather of the updated bf16 parameters, ONE kernel.
// The data-parallel step of a bucket is the two shots of the all-reduce with the optimizer in
// between: rank r reduces slice r of the gradient bucket in the switch, updates the fp32 master
// weights / moments of exactly that slice (optimizer state and traffic are 1/N per rank), and
// multicasts the new bf16 parameters into every rank's parameter buffer.  Same NVLink bytes as a
// plain all-reduce, no separate optimizer pass over all parameters, no post-backward tail.
// (reference: ddp/naive_ddp.py:104-127 all-reduce + a full torch optimizer step on every rank)
// ------------------------------------------------------------------------------------------
struct AdamC {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
  int adamw;
};
__device__ __forceinline__ float adam_step(float p, float g, float& m, float& v, const AdamC& c) {
  if (!c.adamw) g += c.wd * p;
  else p *= (1.f - c.lr * c.wd);
  m = c.beta1 * m + (1.f - c.beta1) * g;
  v = c.beta2 * v + (1.f - c.beta2) * g * g;
  const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  return p - (c.lr / c.bc1) * (m / denom);
}

template <bool kMc, bool kWriteGrad>
__global__ void __launch_bounds__(kCollThreadsMax)
fused_rs_adamw_ag_kernel(const __grid_constant__ PeerPtrs gp, char* gmc,
                         const __grid_constant__ PeerPtrs pp, char* pmc, int rank, int world,
                         FusedAdamLaunch a, int slot_base) {
  AdamC c;
  {
    const float step = a.hyper[0];
    c.lr = a.hyper[1];
    c.beta1 = a.beta1; c.beta2 = a.beta2; c.eps = a.eps; c.wd = a.weight_decay;
    c.bc1 = 1.f - powf(a.beta1, step);
    c.bc2_sqrt = sqrtf(1.f - powf(a.beta2, step));
    c.adamw = a.adamw_mode;
  }
  block_barrier(gp, rank, world, slot_base);          // every rank's gradients of this bucket are final
  const size_t per = (a.n_vec + world - 1) / world;
  const size_t lo = min(per * rank, a.n_vec), hi = min(lo + per, a.n_vec);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  constexpr int kUnroll = 4;
  for (size_t i = lo + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
       i += stride * kUnroll) {
    uint4 g[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx < hi) g[u] = reduce16<kMc, false>(gp, gmc, world, a.grad_offset + idx * 16, a.grad_scale);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx >= hi) continue;
      const size_t li = (idx - lo) * 2;                 // float4 index inside my shard
      float4* mp = reinterpret_cast<float4*>(a.master) + li;
      float4* m1 = reinterpret_cast<float4*>(a.exp_avg) + li;
      float4* m2 = reinterpret_cast<float4*>(a.exp_avg_sq) + li;
      float4 p0 = mp[0], p1 = mp[1], e0 = m1[0], e1 = m1[1], s0 = m2[0], s1 = m2[1];
      const float2 ga = unpack_bf16x2(g[u].x), gb = unpack_bf16x2(g[u].y),
                   gc = unpack_bf16x2(g[u].z), gd = unpack_bf16x2(g[u].w);
      p0.x = adam_step(p0.x, ga.x, e0.x, s0.x, c); p0.y = adam_step(p0.y, ga.y, e0.y, s0.y, c);
      p0.z = adam_step(p0.z, gb.x, e0.z, s0.z, c); p0.w = adam_step(p0.w, gb.y, e0.w, s0.w, c);
      p1.x = adam_step(p1.x, gc.x, e1.x, s1.x, c); p1.y = adam_step(p1.y, gc.y, e1.y, s1.y, c);
      p1.z = adam_step(p1.z, gd.x, e1.z, s1.z, c); p1.w = adam_step(p1.w, gd.y, e1.w, s1.w, c);
      mp[0] = p0; mp[1] = p1; m1[0] = e0; m1[1] = e1; m2[0] = s0; m2[1] = s1;
      uint4 o;
      o.x = pack_bf16x2(p0.x, p0.y); o.y = pack_bf16x2(p0.z, p0.w);
      o.z = pack_bf16x2(p1.x, p1.y); o.w = pack_bf16x2(p1.z, p1.w);
      const size_t poff = a.param_offset + idx * 16;
      const size_t goff = a.grad_offset + idx * 16;
      if constexpr (kMc) {
        multimem_st_v4(pmc + poff, o);
        if constexpr (kWriteGrad) multimem_st_v4(gmc + goff, g[u]);
      } else {
#pragma unroll
        for (int q = 0; q < kApiMaxPeers; ++q)
          if (q < world) {
            st_na_v4(reinterpret_cast<char*>(pp.buf[q]) + poff, o);
            if constexpr (kWriteGrad) st_na_v4(reinterpret_cast<char*>(gp.buf[q]) + goff, g[u]);
          }
      }
    }
  }
  block_barrier(gp, rank, world, slot_base + kBarrierSlotWords);   // new parameters visible everywhere
}

// ------------------------------------------------------------------------------------------
// reduce-scatter: slice `rank` of [world x slice] -> out (local)
// ------------------------------------------------------------------------------------------
template <bool kMc, bool kFp32In>
__global__ void __launch_bounds__(kCollThreadsMax)
reduce_scatter_kernel(const __grid_constant__ PeerPtrs pp, const char* mc, int rank, int world, size_t offset,
                      size_t slice_vec, float scale, void* out, int out_fp32, int accumulate,
                      int slot_base) {
  block_barrier(pp, rank, world, slot_base);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t base = offset + static_cast<size_t>(rank) * slice_vec * 16;
  constexpr int kUnroll = kMc ? 8 : 4;
  for (size_t i0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < slice_vec;
       i0 += stride * kUnroll) {
   uint4 vv[kUnroll];
#pragma unroll
   for (int u = 0; u < kUnroll; ++u) {
     const size_t idx = i0 + u * stride;
     if (idx < slice_vec) vv[u] = reduce16<kMc, kFp32In>(pp, mc, world, base + idx * 16, scale);
   }
#pragma unroll
   for (int u = 0; u < kUnroll; ++u) {
    const size_t i = i0 + u * stride;
    if (i >= slice_vec) break;
    const uint4 v = vv[u];
    if constexpr (kFp32In) {
      float4 f = *reinterpret_cast<const float4*>(&v);
      float4* o = reinterpret_cast<float4*>(out) + i;
      if (accumulate) { float4 t = *o; f.x += t.x; f.y += t.y; f.z += t.z; f.w += t.w; }
      *o = f;
    } else if (out_fp32) {
      float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z),
             d = unpack_bf16x2(v.w);
      float4* o = reinterpret_cast<float4*>(out) + 2 * i;
      float4 f0 = make_float4(a.x, a.y, b.x, b.y), f1 = make_float4(c.x, c.y, d.x, d.y);
      if (accumulate) {
        float4 t0 = o[0], t1 = o[1];
        f0.x += t0.x; f0.y += t0.y; f0.z += t0.z; f0.w += t0.w;
        f1.x += t1.x; f1.y += t1.y; f1.z += t1.z; f1.w += t1.w;
      }
      o[0] = f0; o[1] = f1;
    } else {
      uint4* o = reinterpret_cast<uint4*>(out) + i;
      if (accumulate) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        acc_bf16x8(acc, v); acc_bf16x8(acc, *o);
        *o = pack8(acc, 1.f);
      } else {
        *o = v;
      }
    }
   }
  }
  block_barrier(pp, rank, world, slot_base + kBarrierSlotWords);
}

// ------------------------------------------------------------------------------------------
// all-gather: my slice (local src) -> slot `rank` of the symmetric region on every rank
// ------------------------------------------------------------------------------------------
template <bool kMc>
__global__ void __launch_bounds__(kCollThreadsMax)
all_gather_kernel(const __grid_constant__ PeerPtrs pp, char* mc, int rank, int world, size_t offset, size_t slice_vec,
                  const uint4* src, int slot_base, uint32_t* flag_base_unused) {
  block_barrier(pp, rank, world, slot_base);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t base = offset + static_cast<size_t>(rank) * slice_vec * 16;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slice_vec;
       i += stride) {
    const uint4 v = src ? src[i]
                        : *reinterpret_cast<const uint4*>(
                              reinterpret_cast<const char*>(pp.buf[rank]) + base + i * 16);
    if constexpr (kMc) {
      multimem_st_v4(mc + base + i * 16, v);
    } else {
#pragma unroll
      for (int p = 0; p < kApiMaxPeers; ++p)
        if (p < world && (p != rank || src != nullptr))
          st_na_v4(reinterpret_cast<char*>(pp.buf[p]) + base + i * 16, v);
    }
  }
  block_barrier(pp, rank, world, slot_base + kBarrierSlotWords);
}

// all-gather that publishes a per-source-chunk flag instead of a trailing barrier: consumers
// (the fused all-gather->GEMM producer warp) poll flag[src] >= value on their own pad.
// One counter per source rank is bumped once per *block*; consumer target = value * gridDim.x
// is avoided by letting only the last block to finish (device-wide ticket) publish.
// Runs CONCURRENTLY with the persistent GEMM that consumes it, and may be scheduled after that
// GEMM already occupies every SM: the block must fit next to one GEMM CTA (10 warps x 144 regs,
// 229.6 KB smem), hence 256 threads (16K regs) and ~1 KB of static smem.
constexpr int kPushThreads = 256;
template <bool kMc>
__global__ void __launch_bounds__(kPushThreads)
all_gather_signal_kernel(const __grid_constant__ PeerPtrs pp, char* mc, int rank, int world, size_t offset,
                         size_t slice_vec, const uint4* src, size_t flag_word, uint32_t flag_value,
                         uint32_t* ticket) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t base = offset + static_cast<size_t>(rank) * slice_vec * 16;
  constexpr int kUnroll = 8;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slice_vec;
       i += stride * kUnroll) {
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx < slice_vec) v[u] = ld_nc_v4(src + idx);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t idx = i + u * stride;
      if (idx < slice_vec) {
        if constexpr (kMc) {
          multimem_st_v4(mc + base + idx * 16, v[u]);
        } else {
#pragma unroll
          for (int p = 0; p < kApiMaxPeers; ++p)
            if (p < world) st_na_v4(reinterpret_cast<char*>(pp.buf[p]) + base + idx * 16, v[u]);
        }
      }
    }
  }
  __threadfence_system();      // each thread's own peer / multicast stores reach system scope
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    fence_acq_rel_sys();
    const uint32_t t = atomicAdd(ticket, 1u);
    fence_acq_rel_sys();   // acquire side: order the other blocks' stores before the flag publish
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    // every block's stores are ordered before its ticket increment (fence + atomic); the last
    // block observes all tickets, so a system-scope release store now publishes the whole slice.
    if (threadIdx.x < world) {
      fence_acq_rel_sys();
      st_release_sys(pp.signal[threadIdx.x] + flag_word + rank, flag_value);
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// second stage of fused GEMM->reduce-scatter / GEMM->all-reduce
// ------------------------------------------------------------------------------------------
template <bool kMc>
__global__ void __launch_bounds__(kCollThreadsMax)
rs_reduce_kernel(const __grid_constant__ PeerPtrs pp, char* mc, int rank, int world, RsReduceLaunch r) {
  // wait until every source rank has delivered all of its tiles for my chunk
  if (threadIdx.x < world) {
    const uint32_t* cnt = pp.signal[rank] + r.counter_word_offset + threadIdx.x;
    SpinWatchdog wd;
    while (static_cast<int32_t>(ld_acquire_sys(cnt) - r.counter_target) < 0) {
      wd.tick("reduce-scatter tile counter (rs_reduce)", threadIdx.x, rank);
    }
  }
  __syncthreads();
  const int vec_per_row = r.cols / 8;
  const size_t total = static_cast<size_t>(r.rows) * vec_per_row;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const char* stage = reinterpret_cast<const char*>(pp.buf[rank]) + r.offset;
  const size_t slice_bytes = static_cast<size_t>(r.rows) * r.ld * 2;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += stride) {
    const int row = static_cast<int>(i / vec_per_row);
    const int col = static_cast<int>(i - static_cast<size_t>(row) * vec_per_row) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint4 vals[kApiMaxPeers];
#pragma unroll
    for (int p = 0; p < kApiMaxPeers; ++p)
      if (p < world)
        vals[p] = ld_volatile_v4(stage + p * slice_bytes +
                                 (static_cast<size_t>(row) * r.ld + col) * 2);
#pragma unroll
    for (int p = 0; p < kApiMaxPeers; ++p)
      if (p < world) acc_bf16x8(acc, vals[p]);
    if (r.bias) {
      acc_bf16x8(acc, *reinterpret_cast<const uint4*>(
                          reinterpret_cast<const __nv_bfloat16*>(r.bias) + col));
    }
    if (r.residual) {
      acc_bf16x8(acc, *reinterpret_cast<const uint4*>(
                          reinterpret_cast<const __nv_bfloat16*>(r.residual) +
                          static_cast<size_t>(row) * r.ld_res + col));
    }
    const uint4 o = pack8(acc, 1.f);
    if (!r.broadcast) {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(r.out) +
                                static_cast<size_t>(row) * r.ld_out + col) = o;
    } else {
      const size_t off = r.bcast_offset +
                         ((static_cast<size_t>(rank) * r.rows + row) * r.ld_out + col) * 2;
      if constexpr (kMc) {
        multimem_st_v4(mc + off, o);
      } else {
#pragma unroll
        for (int p = 0; p < kApiMaxPeers; ++p)
          if (p < world) st_na_v4(reinterpret_cast<char*>(pp.buf[p]) + off, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// MoE all-to-all (row granular)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
a2a_scatter_rows_kernel(const __grid_constant__ PeerPtrs pp, size_t offset, const uint4* src, int n_rows, int vec_per_row,
                        const int* dst_rank, const int* dst_row) {
  const int warps_per_block = blockDim.x / 32;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + threadIdx.x / 32; row < n_rows;
       row += gridDim.x * warps_per_block) {
    const int dr = dst_row[row];
    if (dr < 0) continue;
    char* dst = reinterpret_cast<char*>(pp.buf[dst_rank[row]]) + offset +
                static_cast<size_t>(dr) * vec_per_row * 16;
    const uint4* s = src + static_cast<size_t>(row) * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) st_na_v4(dst + v * 16, s[v]);
  }
}

__global__ void __launch_bounds__(256)
a2a_gather_rows_kernel(const __grid_constant__ PeerPtrs pp, size_t offset, uint4* out, int n_rows, int vec_per_row,
                       const int* src_rank, const int* src_row, const float* scale,
                       int accumulate) {
  const int warps_per_block = blockDim.x / 32;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + threadIdx.x / 32; row < n_rows;
       row += gridDim.x * warps_per_block) {
    const int sr = src_row[row];
    uint4* o = out + static_cast<size_t>(row) * vec_per_row;
    if (sr < 0) {
      if (!accumulate)
        for (int v = lane; v < vec_per_row; v += 32) o[v] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const char* s = reinterpret_cast<const char*>(pp.buf[src_rank[row]]) + offset +
                    static_cast<size_t>(sr) * vec_per_row * 16;
    const float w = scale ? scale[row] : 1.f;
    for (int v = lane; v < vec_per_row; v += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      acc_bf16x8(acc, ld_nc_v4(s + v * 16));
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] *= w;
      if (accumulate) acc_bf16x8(acc, o[v]);
      o[v] = pack8(acc, 1.f);
    }
  }
}

__global__ void barrier_only_kernel(const __grid_constant__ PeerPtrs pp, int rank, int world, int slot_base) {
  block_barrier(pp, rank, world, slot_base);
}

PeerPtrs to_pp(const SymmPeers& s) {
  PeerPtrs p;
  for (int i = 0; i < kApiMaxPeers; ++i) {
    p.buf[i] = s.buf[i];
    p.signal[i] = s.signal[i];
  }
  return p;
}

int pick_blocks(size_t n_vec_per_rank, int max_ctas) {
  size_t want = (n_vec_per_rank + kCollThreads * 8 - 1) / (kCollThreads * 8);
  // default cap keeps ~16 K threads in flight whatever the CTA shape
  int cap = max_ctas > 0 ? max_ctas : (16384 / kCollThreads);
  if (cap > kMaxCollBlocks) cap = kMaxCollBlocks;
  if (want < 1) want = 1;
  return static_cast<int>(want < static_cast<size_t>(cap) ? want : cap);
}

uint32_t* g_ticket = nullptr;
uint32_t* ticket_counter() {
  if (!g_ticket) {
    cudaMalloc(&g_ticket, 64 * sizeof(uint32_t));
    cudaMemset(g_ticket, 0, 64 * sizeof(uint32_t));
  }
  return g_ticket;
}

}  // namespace

// Signal pad word layout (per symmetric allocation):
//   [0, 2*kBarrierSlotWords)                       barrier slots of collective kernels (A/B)
//   [2*kBarrierSlotWords, 3*kBarrierSlotWords)     standalone barrier
//   [kUserWordBase, ...)                            chunk flags / tile counters (fused GEMM paths)
constexpr int kStandaloneBarrierBase = 2 * kBarrierSlotWords;

void launch_symm_barrier(const SymmPeers& s, int slot, cudaStream_t stream) {
  barrier_only_kernel<<<1, 32, 0, stream>>>(to_pp(s), s.rank, s.world,
                                            kStandaloneBarrierBase + slot * kApiMaxPeers);
}

void launch_all_reduce(const SymmPeers& s, size_t offset, size_t numel, int dtype, float scale,
                       int algo, int max_ctas, cudaStream_t stream, void* one_shot_scratch) {
  const size_t elem = dtype == 1 ? 4 : 2;
  const size_t n_vec = numel * elem / 16;
  // algo 0 (auto): one-shot up to 256 KiB, two-shot above; 1 forces one-shot (P2P loads),
  // 4 one-shot on multimem, 2 / 3 two-shot P2P / NVLS
  // (measured on 2 x B200, scripts/symm_check.py: 15.8 us vs 17.8 us two-shot vs 20.8 us NCCL at
  //  2 KiB; break-even near 32 KiB -- the auto path switches there, explicit algos go to 256 KiB)
  const bool one_shot = one_shot_scratch != nullptr &&
                        ((algo == 0 && n_vec * 16 <= 32 * 1024) ||
                         ((algo == 1 || algo == 4) && n_vec * 16 <= kOneShotScratchBytes));
  if (one_shot) {
    const bool mc1 = algo == 4 || (algo == 0 && s.mc_buf != nullptr);
    const int threads = 256;
    int blocks = static_cast<int>((n_vec + threads - 1) / threads);
    if (blocks < 1) blocks = 1;
    if (blocks > 32) blocks = 32;
    PeerPtrs pp1 = to_pp(s);
    const char* mcp1 = reinterpret_cast<const char*>(s.mc_buf);
    uint4* scratch = reinterpret_cast<uint4*>(one_shot_scratch);
#define TDP_AR1(MC, F32)                                                                       \
  all_reduce_one_shot_kernel<MC, F32><<<blocks, threads, 0, stream>>>(pp1, mcp1, s.rank, s.world, \
                                                                     offset, n_vec, scale, scratch, 0)
    if (mc1) { if (dtype == 1) TDP_AR1(true, true); else TDP_AR1(true, false); }
    else     { if (dtype == 1) TDP_AR1(false, true); else TDP_AR1(false, false); }
#undef TDP_AR1
    return;
  }
  const bool mc = (algo == 3) || (algo == 0 && s.mc_buf != nullptr);
  const int blocks = pick_blocks(n_vec / (s.world > 0 ? s.world : 1), max_ctas);
  PeerPtrs pp = to_pp(s);
  char* mcp = reinterpret_cast<char*>(s.mc_buf);
  if (mc) {
    if (dtype == 1)
      all_reduce_two_shot_kernel<true, true><<<blocks, kCollThreads, 0, stream>>>(
          pp, mcp, s.rank, s.world, offset, n_vec, scale, 0);
    else
      all_reduce_two_shot_kernel<true, false><<<blocks, kCollThreads, 0, stream>>>(
          pp, mcp, s.rank, s.world, offset, n_vec, scale, 0);
  } else {
    if (dtype == 1)
      all_reduce_two_shot_kernel<false, true><<<blocks, kCollThreads, 0, stream>>>(
          pp, mcp, s.rank, s.world, offset, n_vec, scale, 0);
    else
      all_reduce_two_shot_kernel<false, false><<<blocks, kCollThreads, 0, stream>>>(
          pp, mcp, s.rank, s.world, offset, n_vec, scale, 0);
  }
}

void launch_fused_rs_adamw_ag(const SymmPeers& grad, const SymmPeers& param,
                              const FusedAdamLaunch& a, int use_mc, int write_back_grad,
                              int max_ctas, cudaStream_t stream) {
  const bool mc = use_mc && grad.mc_buf != nullptr && param.mc_buf != nullptr;
  const int world = grad.world > 0 ? grad.world : 1;
  const int blocks = pick_blocks(a.n_vec / world / 2 + 1, max_ctas);
  PeerPtrs gp = to_pp(grad), pp = to_pp(param);
  char* gmc = reinterpret_cast<char*>(grad.mc_buf);
  char* pmc = reinterpret_cast<char*>(param.mc_buf);
#define TDP_FUSED(MC, WG)                                                                  \
  fused_rs_adamw_ag_kernel<MC, WG><<<blocks, kCollThreads, 0, stream>>>(gp, gmc, pp, pmc,   \
                                                                         grad.rank, world, a, 0)
  if (mc) { if (write_back_grad) TDP_FUSED(true, true); else TDP_FUSED(true, false); }
  else    { if (write_back_grad) TDP_FUSED(false, true); else TDP_FUSED(false, false); }
#undef TDP_FUSED
}

void launch_reduce_scatter(const SymmPeers& s, size_t offset, size_t slice_numel, int dtype,
                           float scale, void* out, int out_fp32, int accumulate_out, int use_mc,
                           int max_ctas, cudaStream_t stream) {
  const size_t elem = dtype == 1 ? 4 : 2;
  const size_t slice_vec = slice_numel * elem / 16;
  const bool mc = use_mc && s.mc_buf != nullptr;
  const int blocks = pick_blocks(slice_vec, max_ctas);
  PeerPtrs pp = to_pp(s);
  const char* mcp = reinterpret_cast<const char*>(s.mc_buf);
#define TDP_RS(MC, F32)                                                                      \
  reduce_scatter_kernel<MC, F32><<<blocks, kCollThreads, 0, stream>>>(                        \
      pp, mcp, s.rank, s.world, offset, slice_vec, scale, out, out_fp32, accumulate_out, 0)
  if (mc) { if (dtype == 1) TDP_RS(true, true); else TDP_RS(true, false); }
  else    { if (dtype == 1) TDP_RS(false, true); else TDP_RS(false, false); }
#undef TDP_RS
}

void launch_all_gather(const SymmPeers& s, size_t offset, size_t slice_numel, int elem_bytes,
                       const void* src, int use_mc, int max_ctas, cudaStream_t stream) {
  const size_t slice_vec = slice_numel * elem_bytes / 16;
  const bool mc = use_mc && s.mc_buf != nullptr;
  const int blocks = pick_blocks(slice_vec, max_ctas);
  PeerPtrs pp = to_pp(s);
  char* mcp = reinterpret_cast<char*>(s.mc_buf);
  if (mc)
    all_gather_kernel<true><<<blocks, kCollThreads, 0, stream>>>(
        pp, mcp, s.rank, s.world, offset, slice_vec, reinterpret_cast<const uint4*>(src), 0,
        nullptr);
  else
    all_gather_kernel<false><<<blocks, kCollThreads, 0, stream>>>(
        pp, mcp, s.rank, s.world, offset, slice_vec, reinterpret_cast<const uint4*>(src), 0,
        nullptr);
}

void launch_all_gather_signal(const SymmPeers& s, size_t offset, size_t slice_bytes,
                              const void* src, size_t flag_word_offset, uint32_t flag_value,
                              int use_mc, int max_ctas, cudaStream_t stream) {
  const size_t slice_vec = slice_bytes / 16;
  const bool mc = use_mc && s.mc_buf != nullptr;
  const int blocks = pick_blocks(slice_vec, max_ctas);
  PeerPtrs pp = to_pp(s);
  char* mcp = reinterpret_cast<char*>(s.mc_buf);
  uint32_t* ticket = ticket_counter();
  if (mc)
    all_gather_signal_kernel<true><<<blocks, kPushThreads, 0, stream>>>(
        pp, mcp, s.rank, s.world, offset, slice_vec, reinterpret_cast<const uint4*>(src),
        flag_word_offset, flag_value, ticket);
  else
    all_gather_signal_kernel<false><<<blocks, kPushThreads, 0, stream>>>(
        pp, mcp, s.rank, s.world, offset, slice_vec, reinterpret_cast<const uint4*>(src),
        flag_word_offset, flag_value, ticket);
}

void launch_rs_reduce(const SymmPeers& s, const RsReduceLaunch& r, int use_mc, int max_ctas,
                      cudaStream_t stream) {
  const bool mc = use_mc && s.mc_buf != nullptr;
  const size_t total_vec = static_cast<size_t>(r.rows) * (r.cols / 8);
  int blocks = pick_blocks(total_vec, max_ctas > 0 ? max_ctas : kMaxCollBlocks);
  PeerPtrs pp = to_pp(s);
  char* mcp = reinterpret_cast<char*>(s.mc_buf);
  if (mc) rs_reduce_kernel<true><<<blocks, kCollThreads, 0, stream>>>(pp, mcp, s.rank, s.world, r);
  else    rs_reduce_kernel<false><<<blocks, kCollThreads, 0, stream>>>(pp, mcp, s.rank, s.world, r);
}

void launch_a2a_scatter_rows(const SymmPeers& s, size_t offset, const void* src, int n_rows,
                             int hidden, const int* dst_rank, const int* dst_row,
                             cudaStream_t stream) {
  if (n_rows <= 0) return;
  const int vec_per_row = hidden * 2 / 16;
  int blocks = (n_rows + 7) / 8;
  if (blocks > 296) blocks = 296;
  a2a_scatter_rows_kernel<<<blocks, 256, 0, stream>>>(to_pp(s), offset,
                                                      reinterpret_cast<const uint4*>(src), n_rows,
                                                      vec_per_row, dst_rank, dst_row);
}

void launch_a2a_gather_rows(const SymmPeers& s, size_t offset, void* out, int n_rows, int hidden,
                            const int* src_rank, const int* src_row, const float* scale,
                            int accumulate, cudaStream_t stream) {
  if (n_rows <= 0) return;
  const int vec_per_row = hidden * 2 / 16;
  int blocks = (n_rows + 7) / 8;
  if (blocks > 296) blocks = 296;
  a2a_gather_rows_kernel<<<blocks, 256, 0, stream>>>(to_pp(s), offset,
                                                     reinterpret_cast<uint4*>(out), n_rows,
                                                     vec_per_row, src_rank, src_row, scale,
                                                     accumulate);
}

}  // namespace tdp
