// Flash-attention forward for sm_100a: tcgen05.mma with S and the per-tile P.V product in TMEM,
// TMA-fed K / V ring, two softmax warp-groups ping-ponging on two 128-row query tiles so the
// tensor core works on one tile's QK^T / PV while the other tile's exponentials are computed.
//
//   warp 0          TMA producer  (Q tiles once; K(j), V(j) through a 3-stage ring each)
//   warp 1          MMA issuer    (one elected thread) + TMEM allocation
//   warps 2-5       softmax warp-group 0: query rows q0 .. q0+127      (thread = row = TMEM lane)
//   warps 6-9       softmax warp-group 1: query rows q0+128 .. q0+255
//
//   per KV tile j and warp-group w:
//     S_w = Q_w K_j^T            UMMA 128x128x64   -> TMEM cols [128w, 128w+128)
//     WG w: row max, online-softmax rescale, P = exp2(S*c - m) -> bf16 -> swizzled smem (A operand)
//     O_w' = P_w V_j             UMMA 128x64x128   -> TMEM cols [256+64w, ...)   (not accumulated)
//     WG w: o = (o + O_w') * alpha   in registers (fp32), l likewise
//   epilogue: o / l -> bf16 -> swizzled smem -> TMA store; LSE (natural log) -> global.
//
// Layout contract: q, k, v, o are [B*T, ld] row-major "token matrices" whose row r = b*T + t holds
// all heads of a token (head h at columns col0 + h*64 ...): exactly the packed qkv GEMM output
// (q | k | v along the row) and the [B, T, H*D] attention output -- no permutes, no split copies.
// head_dim = 64, T % 128 == 0.
//
// STATUS: written against the same descriptor / barrier building blocks as the GEMM kernels
// (which are validated on B200); this kernel itself has not run on hardware yet -- it is opt-in
// (TDP_ATTN=native) and checked by scripts/attn_check.py.  (reference: attn.py:40-43 is the
// unfused QK^T / softmax / PV this replaces.)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../common/ptx.cuh"
#include "../common/tdp_api.h"
#include "../common/tmap.h"

namespace tdp {

namespace {

constexpr int kHeadDim = 64;
constexpr int kTileQ = 128;          // query rows per warp-group
constexpr int kTileKV = 128;         // keys per iteration
constexpr int kStagesKV = 3;
constexpr int kAttnThreads = 32 * 10;
constexpr int kTileBytes = kTileQ * kHeadDim * 2;        // 16 KiB: one [128 x 64] bf16 tile
constexpr int kPBytes = kTileQ * kTileKV * 2;            // 32 KiB: P tile, two K-major k-blocks
constexpr uint32_t kTmemColsAttn = 512;                  // S0 | S1 | O0' | O1' (384 used)

struct AttnSmem {
  static constexpr int kQ = 0;                                   // 2 tiles
  static constexpr int kK = kQ + 2 * kTileBytes;                 // kStagesKV tiles
  static constexpr int kV = kK + kStagesKV * kTileBytes;         // kStagesKV tiles
  static constexpr int kP = kV + kStagesKV * kTileBytes;         // 2 P tiles
  static constexpr int kBars = kP + 2 * kPBytes;
  static constexpr int kTotal = kBars + 256;                     // 196 864 B
};

struct AttnParams {
  int B, T, H;
  int causal;
  float scale_log2;        // softmax scale * log2(e)
  int q_col0, k_col0, v_col0, o_col0;   // first column of head 0 in the respective token matrix
  float* lse;              // [B, H, T]
};

TDP_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
TDP_DEVICE float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
TDP_DEVICE void wg_bar_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box {64, 128}
                      const __grid_constant__ CUtensorMap tmap_k,   // box {64, 128}
                      const __grid_constant__ CUtensorMap tmap_v,   // box {64, 128}
                      const __grid_constant__ CUtensorMap tmap_o,   // box {64, 128}
                      const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_q = smem + AttnSmem::kQ;
  uint8_t* smem_k = smem + AttnSmem::kK;
  uint8_t* smem_v = smem + AttnSmem::kV;
  uint8_t* smem_p = smem + AttnSmem::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::kBars);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // kStagesKV
  uint64_t* k_empty = k_full + kStagesKV;
  uint64_t* v_full = k_empty + kStagesKV;
  uint64_t* v_empty = v_full + kStagesKV;
  uint64_t* s_full = v_empty + kStagesKV;        // 2: S_w ready for warp-group w
  uint64_t* p_ready = s_full + 2;                // 2: P_w written (and S_w / O_w' consumed)
  uint64_t* o_full = p_ready + 2;                // 2: O_w' = P_w V ready
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  // heavier (later, for causal) query tiles first
  const int qpair = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qpair * 2 * kTileQ;                       // first query row of this CTA
  const int row_base = b * p.T;                            // row of (b, t = 0) in the token matrices
  // number of KV tiles each warp-group needs
  const int n_all = p.T / kTileKV;
  int n_kv[2];
  n_kv[0] = p.causal ? (q0 / kTileKV + 1) : n_all;
  n_kv[1] = (q0 + kTileQ < p.T) ? (p.causal ? (q0 / kTileKV + 2) : n_all) : 0;
  const int n_max = n_kv[0] > n_kv[1] ? n_kv[0] : n_kv[1];

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStagesKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&p_ready[w], 4);      // one elected lane per warp of the warp-group
      mbar_init(&o_full[w], 1);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<kTmemColsAttn>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      const int n_q = n_kv[1] > 0 ? 2 : 1;
      mbar_expect_tx(q_full, n_q * kTileBytes);
      for (int w = 0; w < n_q; ++w)
        tma_load_2d(&tmap_q, q_full, smem_q + w * kTileBytes, p.q_col0 + h * kHeadDim,
                    row_base + q0 + w * kTileQ);
      for (int j = 0; j < n_max; ++j) {
        const int st = j % kStagesKV;
        const uint32_t ph = (j / kStagesKV) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], kTileBytes);
        tma_load_2d(&tmap_k, &k_full[st], smem_k + st * kTileBytes, p.k_col0 + h * kHeadDim,
                    row_base + j * kTileKV);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], kTileBytes);
        // one [128 keys x 64 d] box: rows of 128 B, i.e. two stacked [64 x 64] boxes -- exactly the
        // MN-major B operand of P.V (one 8 KiB box per 64-key k-block)
        tma_load_2d(&tmap_v, &v_full[st], smem_v + st * kTileBytes, p.v_col0 + h * kHeadDim,
                    row_base + j * kTileKV);
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    // S = Q K^T : A = Q  [128 x 64]  K-major,  B = K [128 keys x 64] K-major   (N = 128)
    // O'= P V   : A = P  [128 x 128] K-major (2 k-blocks), B = V [128 keys x 64] MN-major (N = 64)
    const uint32_t idesc_s = make_idesc_bf16_f32(kTileQ, kTileKV, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16_f32(kTileQ, kHeadDim, 0, 1);
    constexpr uint32_t kUmmaKBytes = 16 * 2;          // K = 16 bf16 along a 128-byte swizzled row
    auto issue_s = [&](int w, int st) {
      const uint32_t sa = smem_u32(smem_q + w * kTileBytes);
      const uint32_t sb = smem_u32(smem_k + st * kTileBytes);
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) {
        const uint64_t da = make_umma_smem_desc_sw128(sa + k * kUmmaKBytes, 0, 1024);
        const uint64_t db = make_umma_smem_desc_sw128(sb + k * kUmmaKBytes, 0, 1024);
        umma_f16_ss(tmem_base + w * kTileKV, da, db, idesc_s, k != 0 ? 1u : 0u);
      }
      umma_commit(&s_full[w]);
    };
    auto issue_pv = [&](int w, int st) {
      const uint32_t sp = smem_u32(smem_p + w * kPBytes);
      const uint32_t sv = smem_u32(smem_v + st * kTileBytes);
#pragma unroll
      for (int kb = 0; kb < kTileKV / 64; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // A: k-block kb of P is a [128 x 64] K-major tile; B: box kb of V, 16 key rows per step
          const uint64_t da = make_umma_smem_desc_sw128(sp + kb * kTileBytes + k * kUmmaKBytes, 0, 1024);
          const uint64_t db = make_umma_smem_desc_sw128(sv + kb * (kTileBytes / 2) + k * 16 * 128,
                                                        64 * 64 * 2, 1024);
          umma_f16_ss(tmem_base + 2 * kTileKV + w * kHeadDim, da, db, idesc_o,
                      (kb | k) != 0 ? 1u : 0u);
        }
      }
      umma_commit(&o_full[w]);
    };

    mbar_wait(q_full, 0);
    if (n_max > 0) {
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        for (int w = 0; w < 2; ++w)
          if (n_kv[w] > 0) issue_s(w, 0);
        umma_commit(&k_empty[0]);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_max; ++j) {
      const int st = j % kStagesKV, st1 = (j + 1) % kStagesKV;
      const uint32_t ph = (j / kStagesKV) & 1, ph1 = ((j + 1) / kStagesKV) & 1;
      mbar_wait(&v_full[st], ph);
      if (j + 1 < n_max) mbar_wait(&k_full[st1], ph1);
      for (int w = 0; w < 2; ++w) {
        if (j >= n_kv[w]) continue;
        mbar_wait(&p_ready[w], j & 1);          // P_w(j) in smem, S_w / O_w' free again
        tc_fence_after();
        if (elect_one()) {
          issue_pv(w, st);
          if (j + 1 < n_kv[w]) issue_s(w, st1);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(&v_empty[st]);              // V(j) consumed once both P.V are complete
        if (j + 1 < n_max) umma_commit(&k_empty[st1]);
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax warp-groups ================================
    const int wg = (warp_idx - 2) >> 2;
    const int quad = warp_idx & 3;
    const int row = quad * 32 + lane;                       // row in the 128-row tile = TMEM lane
    const int n_mine = n_kv[wg];
    const uint32_t t_s = tmem_base + wg * kTileKV + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = tmem_base + 2 * kTileKV + wg * kHeadDim + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* my_p = smem_p + wg * kPBytes;
    const int swz = row & 7;
    float o[kHeadDim];
#pragma unroll
    for (int d = 0; d < kHeadDim; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    auto add_o_tile = [&]() {
#pragma unroll
      for (int c = 0; c < kHeadDim / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_o + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(r[i]);
      }
    };

    for (int j = 0; j < n_mine; ++j) {
      mbar_wait(&s_full[wg], j & 1);
      tc_fence_after();
      const bool diag = p.causal && (j == n_mine - 1);     // keys j*128 + c vs query q0+128wg+row
      // ---- pass A: row max of the raw scores
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kTileKV / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_s + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(r[i]);
          if (!diag || c * 32 + i <= row) mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      // ---- fold the previous tile's P.V into the running output, then rescale to the new max
      if (j > 0) {
        mbar_wait(&o_full[wg], (j - 1) & 1);
        tc_fence_after();
        add_o_tile();
      }
      const float alpha = ex2(m_run - m_new);               // 0 on the first tile (m_run = -inf)
#pragma unroll
      for (int d = 0; d < kHeadDim; ++d) o[d] *= alpha;
      l_run *= alpha;
      // ---- pass B: P = exp2(S * c - m), row sum, bf16 P into the K-major swizzled A tile
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < kTileKV / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_s + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
          float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_new));
          if (diag) {
            if (c * 32 + i > row) p0 = 0.f;
            if (c * 32 + i + 1 > row) p1 = 0.f;
          }
          lsum += p0 + p1;
          pk[i / 2] = pack_bf16x2(p0, p1);
        }
        // columns [32c, 32c+32) = k-block c/2, 16-byte chunks (c%2)*4 .. +3 of this row
        uint8_t* dst = my_p + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;
          *reinterpret_cast<uint4*>(dst + ((chunk ^ swz) << 4)) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      l_run += lsum;
      m_run = m_new;
      // S_w and O_w' are drained, P_w is written: hand all three to the MMA warp
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[wg]);
    }

    if (n_mine > 0) {
      mbar_wait(&o_full[wg], (n_mine - 1) & 1);
      tc_fence_after();
      add_o_tile();
      const float inv_l = 1.f / l_run;
      // ---- O tile -> swizzled staging (the P buffer is free: the last P.V has completed) -> TMA
      uint8_t* dst = my_p + row * 128;
#pragma unroll
      for (int q8 = 0; q8 < kHeadDim / 8; ++q8) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * q8] * inv_l, o[8 * q8 + 1] * inv_l);
        v.y = pack_bf16x2(o[8 * q8 + 2] * inv_l, o[8 * q8 + 3] * inv_l);
        v.z = pack_bf16x2(o[8 * q8 + 4] * inv_l, o[8 * q8 + 5] * inv_l);
        v.w = pack_bf16x2(o[8 * q8 + 6] * inv_l, o[8 * q8 + 7] * inv_l);
        *reinterpret_cast<uint4*>(dst + ((q8 ^ swz) << 4)) = v;
      }
      const int t = q0 + wg * kTileQ + row;
      if (p.lse != nullptr && t < p.T)
        p.lse[(static_cast<size_t>(b) * p.H + h) * p.T + t] =
            (m_run + lg2(l_run)) * 0.6931471805599453f;
      fence_proxy_async_smem();
      wg_bar_sync(wg);
      if (quad == 2 && lane == 0) {       // warps 2 and 6 are the first warps of their groups
        tma_store_2d(&tmap_o, my_p, p.o_col0 + h * kHeadDim, row_base + q0 + wg * kTileQ);
        tma_store_commit();
        tma_store_wait<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemColsAttn>(tmem_base);
  }
}

}  // namespace

int launch_attn_fwd(const AttnFwdLaunch& a, cudaStream_t stream, const char** err) {
  static thread_local char msg[192];
  *err = msg;
  msg[0] = 0;
  if (a.D != kHeadDim || a.T % kTileKV != 0 || a.T <= 0) {
    snprintf(msg, sizeof(msg), "attn_fwd: head_dim must be 64 and T a multiple of 128 (D=%d T=%d)",
             a.D, a.T);
    return -1;
  }
  const uint64_t rows = static_cast<uint64_t>(a.B) * a.T;
  CUtensorMap tq, tk, tv, to;
  const uint64_t width = static_cast<uint64_t>(a.H) * a.D;
  if (!make_tmap_2d(&tq, a.q, width, rows, a.ld_q, 64, kTileQ) ||
      !make_tmap_2d(&tk, a.k, width, rows, a.ld_k, 64, kTileKV) ||
      !make_tmap_2d(&tv, a.v, width, rows, a.ld_v, 64, kTileKV) ||
      !make_tmap_2d(&to, a.o, width, rows, a.ld_o, 64, kTileQ)) {
    snprintf(msg, sizeof(msg), "attn_fwd: cuTensorMapEncodeTiled failed");
    return -2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) {
      snprintf(msg, sizeof(msg), "attn_fwd: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  AttnParams p;
  p.B = a.B; p.T = a.T; p.H = a.H;
  p.causal = a.causal;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0; p.o_col0 = 0;   // bases already point at head 0
  p.lse = a.lse;
  dim3 grid((a.T + 2 * kTileQ - 1) / (2 * kTileQ), a.H, a.B);
  attn_fwd_sm100_kernel<<<grid, kAttnThreads, AttnSmem::kTotal, stream>>>(tq, tk, tv, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(msg, sizeof(msg), "attn_fwd launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace tdp
