// Flash-attention backward, key/value-gradient half, for sm_100a (companion of
// attn_bwd_dq_sm100.cu, which runs first and produces dQ and delta; same layout contract as the
// forward kernel).
//
// One CTA owns one 128-key tile (b, h, j) and walks the query tiles i that see it:
//     S   = Q_i K_j^T                 UMMA 128x128x64  -> TMEM [  0,128)
//     dP  = dO_i V_j^T                UMMA 128x128x64  -> TMEM [128,256)
//     two softmax warp-groups (thread = query row, warp-group = 64-key half of the tile):
//         P = exp2(S c - lse),  dS = P (dP - delta)   -> bf16 P and dS tiles in swizzled smem
//                                                        (double buffered)
//     dV += P^T  dO_i                 UMMA 128x64x128  -> TMEM [256,320)   accumulated over i
//     dK += dS^T Q_i                  UMMA 128x64x128  -> TMEM [320,384)   accumulated over i
//   end: dV, scale * dK -> bf16 -> swizzled smem -> TMA store (into the k / v windows of the packed
//   dqkv gradient).
//
// Every smem tile is used through two descriptor views without ever being transposed:
//   Q_i / dO_i / K_j  [128 rows x 64]   as K-major A/B (rows = M/N) and as MN-major B (rows = K)
//   P / dS  stored as [key half][q half][64 q][64 keys] boxes, read as MN-major A for P^T dO /
//   dS^T Q (LBO = 16 KiB between key halves).
//
//   warps 0-3 / 4-7  softmax warp-groups (key half 0 / 1)    warp 8  TMA producer
//   warp 9  MMA issuer + TMEM allocation                      warps 10-11 idle (setmaxnreg group)
//
// Validated on B200 by scripts/attn_check.py and tests/test_gpu_kernels.py.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../common/ptx.cuh"
#include "../common/tdp_api.h"
#include "../common/tmap.h"
#include "attn_common.cuh"

namespace tdp {

namespace {

constexpr int kD = 64;
constexpr int kT = 128;                       // tile edge (queries and keys)
constexpr int kTile = kT * kD * 2;            // 16 KiB  [128 x 64] bf16
constexpr int kPTile = kT * kT * 2;           // 32 KiB  [128 x 128] bf16
constexpr int kQStages = 2;
constexpr int kBwdThreads = 32 * 12;
constexpr uint32_t kBwdTmemCols = 512;

struct BwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;                    // kQStages x (Q_i | dO_i)
  static constexpr int kPds = kQ + kQStages * 2 * kTile;   // 2 buffers x (P | dS)
  static constexpr int kBars = kPds + 2 * 2 * kPTile;
  static constexpr int kTotal = kBars + 256;               // 229 632 B
};

struct BwdParams {
  int B, T, H;
  int causal;
  float scale, scale_log2;
  const float* lse;        // [B, H, T] natural log
  const float* delta;      // [B, H, T] rowsum(dO * O)  (written by the dQ kernel)
};

using namespace attn;

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q,    // box {64,128}
                      const __grid_constant__ CUtensorMap tmap_k,    // box {64,128}
                      const __grid_constant__ CUtensorMap tmap_v,    // box {64,128}
                      const __grid_constant__ CUtensorMap tmap_do,   // box {64,128}
                      const __grid_constant__ CUtensorMap tmap_dk,   // box {64,128}
                      const __grid_constant__ CUtensorMap tmap_dv,   // box {64,128}
                      const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_k = smem + BwdSmem::kK;
  uint8_t* smem_v = smem + BwdSmem::kV;
  uint8_t* smem_q = smem + BwdSmem::kQ;          // stage s: Q at +s*32K, dO at +s*32K+16K
  uint8_t* smem_pds = smem + BwdSmem::kPds;      // buffer u: P at +u*64K, dS at +u*64K+32K
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBars);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                   // kQStages
  uint64_t* q_empty = q_full + kQStages;         // kQStages
  uint64_t* s_full = q_empty + kQStages;         // S_i and dP_i ready
  uint64_t* p_ready = s_full + 1;                // 2 (per P/dS buffer): written, S/dP consumed
  uint64_t* pds_free = p_ready + 2;              // 2: the MMAs reading buffer u have completed
  uint64_t* sdp_free = pds_free + 2;             // S_i / dP_i are in registers (TMEM reusable)
  uint64_t* dkv_full = sdp_free + 1;             // all of dK / dV accumulated
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(dkv_full + 1);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;       // key tile, head, batch
  const int n_tiles = p.T / kT;
  const int i0 = p.causal ? j : 0;                                // first query tile that sees j
  const int n_iter = n_tiles - i0;
  const int row_base = b * p.T;

  if (warp_idx == 8 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    for (int u = 0; u < 2; ++u) {
      mbar_init(&p_ready[u], 8);               // one elected lane per softmax warp (2 groups)
      mbar_init(&pds_free[u], 1);
    }
    mbar_init(sdp_free, 8);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  } else if (warp_idx == 9) {
    tmem_alloc<kBwdTmemCols>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320;

  if (warp_idx >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp_idx == 8) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * kTile);
      tma_load_2d(&tmap_k, kv_full, smem_k, h * kD, row_base + j * kT);
      tma_load_2d(&tmap_v, kv_full, smem_v, h * kD, row_base + j * kT);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kQStages;
        const uint32_t ph = (it / kQStages) & 1;
        mbar_wait(&q_empty[s], ph ^ 1);
        mbar_expect_tx(&q_full[s], 2 * kTile);
        tma_load_2d(&tmap_q, &q_full[s], smem_q + s * 2 * kTile, h * kD, row_base + (i0 + it) * kT);
        tma_load_2d(&tmap_do, &q_full[s], smem_q + s * 2 * kTile + kTile, h * kD,
                    row_base + (i0 + it) * kT);
      }
    }
  } else if (warp_idx == 9) {
    // ================================ MMA issuer ================================
    const uint32_t idesc_s = make_idesc_bf16_f32(kT, kT, 0, 0);     // A K-major, B K-major, N=128
    const uint32_t idesc_t = make_idesc_bf16_f32(kT, kD, 1, 1);     // A^T (MN-major), B MN-major
    const uint32_t idesc_q = make_idesc_bf16_f32(kT, kD, 0, 1);     // A K-major, B MN-major
    const uint32_t sk = smem_u32(smem_k), sv = smem_u32(smem_v);

    // C[128 x 128] = A[128 x 64] B[128 x 64]^T, both K-major tiles
    auto mma_nt = [&](uint32_t col, uint32_t sa, uint32_t sb) {
#pragma unroll
      for (int k = 0; k < kD / 16; ++k)
        umma_f16_ss(tmem_base + col, make_umma_smem_desc_sw128(sa + k * 32, 0, 1024),
                    make_umma_smem_desc_sw128(sb + k * 32, 0, 1024), idesc_s, k != 0 ? 1u : 0u);
    };
    // C[128 keys x 64] (+)= X^T[128 keys x 128 q] Y[128 q x 64]:  X = P or dS buffer (MN-major A,
    // LBO 16 KiB between key halves), Y = dO_i or Q_i tile read as MN-major B (64-row boxes)
    auto mma_tn = [&](uint32_t col, uint32_t sx, uint32_t sy, bool accumulate) {
#pragma unroll
      for (int qb = 0; qb < 2; ++qb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_umma_smem_desc_sw128(sx + qb * (kTile / 2) + k * 16 * 128, kTile, 1024);
          const uint64_t db = make_umma_smem_desc_sw128(sy + qb * (kTile / 2) + k * 16 * 128, kTile / 2, 1024);
          umma_f16_ss(tmem_base + col, da, db, idesc_t, (accumulate || (qb | k) != 0) ? 1u : 0u);
        }
      }
    };
    mbar_wait(kv_full, 0);
    if (n_iter > 0) {
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        mma_nt(kColS, smem_u32(smem_q), sk);                    // S_0
        mma_nt(kColDP, smem_u32(smem_q + kTile), sv);           // dP_0
        umma_commit(s_full);
      }
      __syncwarp();
    }
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kQStages, u = it & 1;
      const uint32_t sq = smem_u32(smem_q + s * 2 * kTile), sdo = sq + kTile;
      const uint32_t sp = smem_u32(smem_pds + u * 2 * kPTile), sds = sp + kPTile;
      // S_i / dP_i have been read into registers: produce the next pair right away, so the
      // softmax groups never wait for the tensor core (their registers are the second buffer)
      mbar_wait(sdp_free, it & 1);
      if (it + 1 < n_iter) {
        const int s1 = (it + 1) % kQStages;
        mbar_wait(&q_full[s1], ((it + 1) / kQStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_nt(kColS, smem_u32(smem_q + s1 * 2 * kTile), sk);
          mma_nt(kColDP, smem_u32(smem_q + s1 * 2 * kTile + kTile), sv);
          umma_commit(s_full);
        }
        __syncwarp();
      }
      mbar_wait(&p_ready[u], (it >> 1) & 1);                    // P_i, dS_i written
      tc_fence_after();
      if (elect_one()) {
        mma_tn(kColDV, sp, sdo, it > 0);                        // dV += P^T dO_i
        mma_tn(kColDK, sds, sq, it > 0);                        // dK += dS^T Q_i
        umma_commit(&pds_free[u]);
        umma_commit(&q_empty[s]);
        if (it == n_iter - 1) umma_commit(dkv_full);
      }
      __syncwarp();
    }
   }
  } else {
    // ================================ softmax warp-groups ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int wg = warp_idx >> 2;                      // key half handled by this warp-group
    const int quad = warp_idx & 3;
    const int row = quad * 32 + lane;                  // query row in the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int swz = row & 7;
    const float* lse_bh = p.lse + (static_cast<size_t>(b) * p.H + h) * p.T;
    const float* delta_bh = p.delta + (static_cast<size_t>(b) * p.H + h) * p.T;
    for (int it = 0; it < n_iter; ++it) {
      const int i = i0 + it, u = it & 1;
      const float lse2 = lse_bh[i * kT + row] * 1.4426950408889634f;
      const float dlt = delta_bh[i * kT + row];
      uint8_t* pbuf = smem_pds + u * 2 * kPTile;
      uint8_t* dsbuf = pbuf + kPTile;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t rs[64], rp[64];
      tmem_ld_x32_at(tmem_base + kColS + wg * 64 + lane_off, rs);
      tmem_ld_x32_at(tmem_base + kColS + wg * 64 + 32 + lane_off, rs + 32);
      tmem_ld_x32_at(tmem_base + kColDP + wg * 64 + lane_off, rp);
      tmem_ld_x32_at(tmem_base + kColDP + wg * 64 + 32 + lane_off, rp + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      // causal diagonal tile (i == j): key wg*64 + e is visible to query `row` iff <= row
      const int lim = (p.causal && i == j) ? (row - wg * 64) : 64;
      uint32_t pp[32], dd[32];
#pragma unroll
      for (int e = 0; e < 64; e += 2) {
        float p0 = ex2(fmaf(__uint_as_float(rs[e]), p.scale_log2, -lse2));
        float p1 = ex2(fmaf(__uint_as_float(rs[e + 1]), p.scale_log2, -lse2));
        if (e > lim) p0 = 0.f;
        if (e + 1 > lim) p1 = 0.f;
        pp[e / 2] = pack_bf16x2(p0, p1);
        dd[e / 2] = pack_bf16x2(p0 * (__uint_as_float(rp[e]) - dlt),
                                p1 * (__uint_as_float(rp[e + 1]) - dlt));
      }
      // S / dP of this tile are in registers; the buffer may still be read by the MMAs of it-2
      if (it >= 2) mbar_wait(&pds_free[u], ((it - 2) >> 1) & 1);
      // key half wg; box (key half, q half) = rows of 128 B
      const int off = wg * kTile + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int chunk = (c ^ swz) << 4;
        *reinterpret_cast<uint4*>(pbuf + off + chunk) =
            make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        *reinterpret_cast<uint4*>(dsbuf + off + chunk) =
            make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[u]);
    }
    // ---- accumulators -> bf16 -> swizzled staging (the Q / dO stages are idle now) -> TMA:
    //      warp-group 0 writes dV, warp-group 1 writes scale * dK (thread = key row)
    if (n_iter > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
      uint8_t* stage = smem_q + wg * kTile;           // 16 KiB each
      const uint32_t col = wg == 0 ? kColDV : kColDK;
      const float mul = wg == 0 ? 1.f : p.scale;
      uint32_t r[kD];
      tmem_ld_x32_at(tmem_base + col + lane_off, r);
      tmem_ld_x32_at(tmem_base + col + 32 + lane_off, r + 32);
      tmem_ld_wait();
      uint8_t* dst = stage + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[8 * c]) * mul, __uint_as_float(r[8 * c + 1]) * mul);
        v.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul);
        v.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul);
        v.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + ((c ^ swz) << 4)) = v;
      }
      fence_proxy_async_smem();
      wg_bar_sync(wg);
      if (quad == 0 && lane == 0) {
        tma_store_2d(wg == 0 ? &tmap_dv : &tmap_dk, stage, h * kD, row_base + j * kT);
        tma_store_commit();
        tma_store_wait<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) {
    tc_fence_after();
    tmem_dealloc<kBwdTmemCols>(tmem_base);
  }
}

}  // namespace

int launch_attn_bwd_dkv(const AttnBwdLaunch& a, cudaStream_t stream, const char** err) {
  static thread_local char msg[192];
  *err = msg;
  msg[0] = 0;
  if (a.D != kD || a.T % kT != 0 || a.T <= 0) {
    snprintf(msg, sizeof(msg), "attn_bwd: head_dim must be 64 and T a multiple of 128 (D=%d T=%d)",
             a.D, a.T);
    return -1;
  }
  const uint64_t rows = static_cast<uint64_t>(a.B) * a.T;
  const uint64_t width = static_cast<uint64_t>(a.H) * a.D;
  CUtensorMap tq, tk, tv, tdo, tdk, tdv;
  if (!make_tmap_2d(&tq, a.q, width, rows, a.ld_q, 64, kT) ||
      !make_tmap_2d(&tk, a.k, width, rows, a.ld_k, 64, kT) ||
      !make_tmap_2d(&tv, a.v, width, rows, a.ld_v, 64, kT) ||
      !make_tmap_2d(&tdo, a.d_o, width, rows, a.ld_do, 64, kT) ||
      !make_tmap_2d(&tdk, a.dk, width, rows, a.ld_dk, 64, kT) ||
      !make_tmap_2d(&tdv, a.dv, width, rows, a.ld_dv, 64, kT)) {
    snprintf(msg, sizeof(msg), "attn_bwd: cuTensorMapEncodeTiled failed");
    return -2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_sm100_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal);
    if (e != cudaSuccess) {
      snprintf(msg, sizeof(msg), "attn_bwd: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  BwdParams p;
  p.B = a.B; p.T = a.T; p.H = a.H;
  p.causal = a.causal;
  p.scale = a.scale;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.lse = a.lse; p.delta = a.delta;
  dim3 grid(a.T / kT, a.H, a.B);
  attn_bwd_sm100_kernel<<<grid, kBwdThreads, BwdSmem::kTotal, stream>>>(tq, tk, tv, tdo, tdk, tdv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(msg, sizeof(msg), "attn_bwd launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int launch_attn_bwd(const AttnBwdLaunch& a, cudaStream_t stream, const char** err) {
  const int rc = launch_attn_bwd_dq(a, stream, err);
  if (rc != 0) return rc;
  return launch_attn_bwd_dkv(a, stream, err);
}

}  // namespace tdp
