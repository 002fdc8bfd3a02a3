// Flash-attention backward, query-gradient half, for sm_100a (same layout contract as the forward
// kernel).  The backward is split into two kernels that each keep their output stationary in TMEM
// -- this one (dQ, plus delta = rowsum(dO . O)) and attn_bwd_sm100.cu (dK, dV) -- instead of one
// kernel that adds dQ tiles into global memory with fp32 atomics: 7 instead of 5 tile GEMMs, but no
// atomics, no accumulator zero-fill, no fp32 -> bf16 conversion pass, bitwise deterministic.
//
//   work item = (b, h, 256 query rows), persistent CTAs, static heavy-first snake schedule
//   (identical machinery to attn_fwd_sm100.cu).  Keys stream through in 64-key sub-tiles so that
//   both warp-groups' S / dP slices *and* both dQ accumulators fit TMEM:
//
//   warps 0-3 / 4-7   softmax warp-groups (thread = query row = TMEM lane), one 128-row tile each
//   warp 8            TMA producer: Q_w, dO_w, O_w per item; K'(j), V'(j) rings (4 stages x 8 KiB)
//   warp 9            MMA issuer + TMEM allocation           warps 10-11 idle (setmaxnreg group)
//
//   item start, WG w:  delta = rowsum(dO_w . O_w) from the two smem tiles (thread = row), -> global
//   per sub-tile j, WG w:
//     S'  = Q_w  K'_j^T          UMMA 128x64x64   -> TMEM cols [64w, 64w+64)
//     dP' = dO_w V'_j^T          UMMA 128x64x64   -> TMEM cols [128+64w, ...)
//     WG w: P = exp2(S' c - lse), dS = P (dP' - delta) -> bf16 -> swizzled smem (A operand)
//     dQ_w += dS K'_j            UMMA 128x64x64   -> TMEM cols [256+64w, ...)   accumulated over j
//   item end: dQ_w * scale -> bf16 -> swizzled smem -> TMA store (straight into the q window of
//   the packed dqkv gradient).
//
// (reference: the autograd backward of parallel/tensor_parallel/attn.py:40-43.)
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

using namespace attn;

constexpr int kD = 64;
constexpr int kTQ = 128;                     // query rows per warp-group
constexpr int kSub = 64;                     // keys per step
constexpr int kStages = 4;
constexpr int kThreads = 32 * 12;
constexpr int kQTile = kTQ * kD * 2;         // 16 KiB
constexpr int kKTile = kSub * kD * 2;        // 8 KiB
constexpr uint32_t kTmemCols = 512;          // S'0 S'1 | dP'0 dP'1 | dQ0 dQ1  (384 used)

struct DqSmem {
  static constexpr int kQ = 0;                              // 2 tiles
  static constexpr int kDO = kQ + 2 * kQTile;               // 2 tiles
  static constexpr int kDS = kDO + 2 * kQTile;              // 2 tiles: O_w (item start), dS_w, dQ_w staging
  static constexpr int kK = kDS + 2 * kQTile;               // kStages sub-tiles
  static constexpr int kV = kK + kStages * kKTile;
  static constexpr int kBars = kV + kStages * kKTile;
  static constexpr int kTotal = kBars + 256;                // 164 096 B
};

struct DqParams {
  int B, T, H;
  int causal;
  float scale, scale_log2;
  int n_qp, n_items;
  const float* lse;        // [B, H, T] natural log (forward)
  float* delta;            // [B, H, T] out: rowsum(dO * O)
};

struct Item {
  int b, h, q0;
  int n_sub[2];
  int n_max;
};
TDP_DEVICE bool get_item(const DqParams& p, int round, Item& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = round * G + ((round & 1) ? (G - 1 - c) : c);
  if (idx >= p.n_items) return false;
  const int bh_count = p.B * p.H;
  const int qp = p.n_qp - 1 - idx / bh_count;
  const int bh = idx - (idx / bh_count) * bh_count;
  it.b = bh / p.H;
  it.h = bh - it.b * p.H;
  it.q0 = qp * 2 * kTQ;
  const int n_all = p.T / kSub;
  it.n_sub[0] = p.causal ? (it.q0 / kSub + 2) : n_all;
  it.n_sub[1] = (it.q0 + kTQ < p.T) ? (p.causal ? (it.q0 / kSub + 4) : n_all) : 0;
  it.n_max = it.n_sub[0] > it.n_sub[1] ? it.n_sub[0] : it.n_sub[1];
  return true;
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q,    // box {64, 128}
                         const __grid_constant__ CUtensorMap tmap_do,   // box {64, 128}
                         const __grid_constant__ CUtensorMap tmap_o,    // box {64, 128}
                         const __grid_constant__ CUtensorMap tmap_k,    // box {64, 64}
                         const __grid_constant__ CUtensorMap tmap_v,    // box {64, 64}
                         const __grid_constant__ CUtensorMap tmap_dq,   // box {64, 128}
                         const DqParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_q = smem + DqSmem::kQ;
  uint8_t* smem_do = smem + DqSmem::kDO;
  uint8_t* smem_ds = smem + DqSmem::kDS;
  uint8_t* smem_k = smem + DqSmem::kK;
  uint8_t* smem_v = smem + DqSmem::kV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::kBars);
  uint64_t* q_full = bars;                       // Q_w, dO_w, O_w of the item landed
  uint64_t* q_empty = bars + 1;                  // every S' / dP' MMA of the item has completed
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;          // 2: S'_w and dP'_w ready
  uint64_t* s_free = s_full + 2;                 // 2: S'_w / dP'_w are in registers (TMEM reusable)
  uint64_t* p_ready = s_free + 2;                // 2: dS_w written
  uint64_t* dq_full = p_ready + 2;               // 2: dQ_w += dS_w K' (one completion per sub-tile)
  uint64_t* stage_free = dq_full + 2;            // 2: the dQ_w store has read its staging tile
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(stage_free + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 8 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_o);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&s_free[w], 4);
      mbar_init(&p_ready[w], 4);
      mbar_init(&dq_full[w], 1);
      mbar_init(&stage_free[w], 1);
    }
    fence_barrier_init();
  } else if (warp_idx == 9) {
    tmem_alloc<kTmemCols>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256;

  if (warp_idx >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp_idx == 8) {
      // ================================ TMA producer ================================
      if (elect_one()) {
        uint32_t g = 0;
        uint32_t n_store[2] = {0u, 0u};        // items in which warp-group w stored a dQ tile
        Item it;
        for (int round = 0; get_item(p, round, it); ++round) {
          const int row_base = it.b * p.T;
          if (round > 0) mbar_wait(q_empty, (round - 1) & 1);
          const int n_q = it.n_sub[1] > 0 ? 2 : 1;
          for (int w = 0; w < n_q; ++w)          // O_w lands in the tile the last dQ_w left through
            if (n_store[w] > 0) mbar_wait(&stage_free[w], (n_store[w] - 1) & 1);
          mbar_expect_tx(q_full, n_q * 3 * kQTile);
          for (int w = 0; w < n_q; ++w) {
            const int r0 = row_base + it.q0 + w * kTQ;
            tma_load_2d(&tmap_q, q_full, smem_q + w * kQTile, it.h * kD, r0);
            tma_load_2d(&tmap_do, q_full, smem_do + w * kQTile, it.h * kD, r0);
            tma_load_2d(&tmap_o, q_full, smem_ds + w * kQTile, it.h * kD, r0);
            ++n_store[w];
          }
          for (int j = 0; j < it.n_max; ++j, ++g) {
            const int st = g % kStages;
            const uint32_t ph = (g / kStages) & 1;
            mbar_wait(&k_empty[st], ph ^ 1);
            mbar_expect_tx(&k_full[st], kKTile);
            tma_load_2d(&tmap_k, &k_full[st], smem_k + st * kKTile, it.h * kD, row_base + j * kSub);
            mbar_wait(&v_empty[st], ph ^ 1);
            mbar_expect_tx(&v_full[st], kKTile);
            tma_load_2d(&tmap_v, &v_full[st], smem_v + st * kKTile, it.h * kD, row_base + j * kSub);
          }
        }
      }
    } else if (warp_idx == 9) {
      // ================================ MMA issuer ================================
      const uint32_t idesc_s = make_idesc_bf16_f32(kTQ, kSub, 0, 0);   // A, B K-major, N = 64
      const uint32_t idesc_q = make_idesc_bf16_f32(kTQ, kD, 0, 1);     // A K-major, B MN-major
      // C[128 x 64] = A[128 x 64 d] B[64 keys x 64 d]^T (both K-major)
      auto mma_nt = [&](uint32_t col, uint32_t sa, uint32_t sb) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_f16_ss(tmem_base + col, make_umma_smem_desc_sw128(sa + k * 32, 0, 1024),
                      make_umma_smem_desc_sw128(sb + k * 32, 0, 1024), idesc_s, k != 0 ? 1u : 0u);
      };
      // dQ[128 x 64 d] (+)= dS[128 x 64 keys] K'[64 keys x 64 d]: A K-major, B MN-major (rows = K)
      auto mma_dq = [&](uint32_t col, uint32_t sa, uint32_t sb, bool accumulate) {
#pragma unroll
        for (int k = 0; k < kSub / 16; ++k)
          umma_f16_ss(tmem_base + col, make_umma_smem_desc_sw128(sa + k * 32, 0, 1024),
                      make_umma_smem_desc_sw128(sb + k * 16 * 128, kKTile, 1024), idesc_q,
                      (accumulate || k != 0) ? 1u : 0u);
      };
      auto issue_sdp = [&](int w, int st) {
        mma_nt(kColS + w * kSub, smem_u32(smem_q + w * kQTile), smem_u32(smem_k + st * kKTile));
        mma_nt(kColDP + w * kSub, smem_u32(smem_do + w * kQTile), smem_u32(smem_v + st * kKTile));
        umma_commit(&s_full[w]);
      };

      // Event-driven issue loop (one thread), same scheme as the forward kernel:
      //   s_free[w]  (S'_w, dP'_w(j) are in registers)  -> issue S'_w, dP'_w(j+1)
      //   p_ready[w] (dS_w(j) is in smem)                -> issue dQ_w += dS_w(j) K'(j)
      if (elect_one()) {
        uint32_t g = 0;
        uint32_t a_cnt[2] = {0u, 0u}, b_cnt[2] = {0u, 0u};
        Item it;
        for (int round = 0; get_item(p, round, it); ++round) {
          int a_loc[2] = {0, 0}, b_loc[2] = {0, 0};
          int s_iss[2] = {0, 0};
          int rk = 0, rv = 0;
          bool q_released = false;
          auto release = [&]() {
            // K'(jj): read by S'(jj) and by dQ(jj) of both groups; V'(jj): by dP'(jj)
            while (rk < it.n_max && (b_loc[0] > rk || rk >= it.n_sub[0]) &&
                   (b_loc[1] > rk || rk >= it.n_sub[1])) {
              umma_commit(&k_empty[(g + rk) % kStages]);
              ++rk;
            }
            while (rv < it.n_max && (s_iss[0] > rv || rv >= it.n_sub[0]) &&
                   (s_iss[1] > rv || rv >= it.n_sub[1])) {
              umma_commit(&v_empty[(g + rv) % kStages]);
              ++rv;
            }
            if (!q_released && s_iss[0] >= it.n_sub[0] && s_iss[1] >= it.n_sub[1]) {
              umma_commit(q_empty);
              q_released = true;
            }
          };
          mbar_wait(q_full, round & 1);
          mbar_wait(&k_full[g % kStages], (g / kStages) & 1);
          mbar_wait(&v_full[g % kStages], (g / kStages) & 1);
          tc_fence_after();
          for (int w = 0; w < 2; ++w)
            if (it.n_sub[w] > 0) { issue_sdp(w, g % kStages); s_iss[w] = 1; }
          release();
          while (a_loc[0] < it.n_sub[0] || b_loc[0] < it.n_sub[0] || a_loc[1] < it.n_sub[1] ||
                 b_loc[1] < it.n_sub[1]) {
            for (int w = 0; w < 2; ++w) {
              if (a_loc[w] < it.n_sub[w] && mbar_test_wait(&s_free[w], a_cnt[w] & 1)) {
                const int j = a_loc[w];
                const uint32_t gj = g + j + 1;
                if (j + 1 >= it.n_sub[w]) {
                  ++a_cnt[w];
                  ++a_loc[w];
                } else if (mbar_test_wait(&k_full[gj % kStages], (gj / kStages) & 1) &&
                           mbar_test_wait(&v_full[gj % kStages], (gj / kStages) & 1)) {
                  ++a_cnt[w];
                  ++a_loc[w];
                  tc_fence_after();
                  issue_sdp(w, gj % kStages);
                  s_iss[w] = j + 2;
                  release();
                }
              }
              if (b_loc[w] < it.n_sub[w] && mbar_test_wait(&p_ready[w], b_cnt[w] & 1)) {
                ++b_cnt[w];
                const int j = b_loc[w];
                tc_fence_after();
                // K'(j) is resident: S'_w(j) was computed from it and it is released below
                mma_dq(kColDQ + w * kD, smem_u32(smem_ds + w * kQTile),
                       smem_u32(smem_k + ((g + j) % kStages) * kKTile), j > 0);
                umma_commit(&dq_full[w]);
                b_loc[w] = j + 1;
                release();
              }
            }
          }
          g += it.n_max;
        }
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax warp-groups ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int wg = warp_idx >> 2;
    const int quad = warp_idx & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + kColS + wg * kSub + lane_off;
    const uint32_t t_dp = tmem_base + kColDP + wg * kSub + lane_off;
    const uint32_t t_dq = tmem_base + kColDQ + wg * kD + lane_off;
    uint8_t* my_ds = smem_ds + wg * kQTile;
    const uint8_t* my_do = smem_do + wg * kQTile;
    const int swz = row & 7;
    uint32_t cnt_s = 0, cnt_dq = 0;
    Item it;
    for (int round = 0; get_item(p, round, it); ++round) {
      const int n_mine = it.n_sub[wg];
      if (n_mine == 0) continue;
      const int t = it.q0 + wg * kTQ + row;                   // my query position
      const size_t stat_idx = (static_cast<size_t>(it.b) * p.H + it.h) * p.T + t;
      const float lse2 = p.lse[stat_idx] * 1.4426950408889634f;
      // ---- delta = rowsum(dO . O): both tiles are in smem (O sits in my dS tile), thread = row
      mbar_wait(q_full, round & 1);
      float dlt = 0.f;
      {
        const uint8_t* prow_do = my_do + row * 128;
        const uint8_t* prow_o = my_ds + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = *reinterpret_cast<const uint4*>(prow_do + ((c ^ swz) << 4));
          const uint4 b = *reinterpret_cast<const uint4*>(prow_o + ((c ^ swz) << 4));
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = unpack_bf16x2(aw[e]), fb = unpack_bf16x2(bw[e]);
            dlt = fmaf(fa.x, fb.x, dlt);
            dlt = fmaf(fa.y, fb.y, dlt);
          }
        }
      }
      p.delta[stat_idx] = dlt;
      for (int j = 0; j < n_mine; ++j) {
        mbar_wait(&s_full[wg], cnt_s & 1);
        ++cnt_s;
        tc_fence_after();
        uint32_t rs[kSub], rp[kSub];
        tmem_ld_x32_at(t_s, rs);
        tmem_ld_x32_at(t_s + 32, rs + 32);
        tmem_ld_x32_at(t_dp, rp);
        tmem_ld_x32_at(t_dp + 32, rp + 32);
        tmem_ld_wait();
        // both slices are in registers: let the tensor core produce the next pair right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);
        // keys j*64 + c are visible to query t iff j*64 + c <= t (only the last sub-tiles cut)
        const int lim = p.causal ? (t - j * kSub) : kSub;     // columns c <= lim are kept
        uint32_t dd[kSub / 2];
#pragma unroll
        for (int e = 0; e < kSub; e += 2) {
          float p0 = ex2(fmaf(__uint_as_float(rs[e]), p.scale_log2, -lse2));
          float p1 = ex2(fmaf(__uint_as_float(rs[e + 1]), p.scale_log2, -lse2));
          if (e > lim) p0 = 0.f;
          if (e + 1 > lim) p1 = 0.f;
          dd[e / 2] = pack_bf16x2(p0 * (__uint_as_float(rp[e]) - dlt),
                                  p1 * (__uint_as_float(rp[e + 1]) - dlt));
        }
        if (j > 0) {                       // dQ(j-1) has finished reading my dS tile
          mbar_wait(&dq_full[wg], cnt_dq & 1);
          ++cnt_dq;
        }
        uint8_t* dst = my_ds + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dst + ((c ^ swz) << 4)) =
              make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[wg]);
      }
      // ---- dQ_w * scale -> bf16 -> staging (my dS tile: its last reader has completed) -> TMA
      mbar_wait(&dq_full[wg], cnt_dq & 1);
      ++cnt_dq;
      tc_fence_after();
      {
        uint32_t o[kD];
        tmem_ld_x32_at(t_dq, o);
        tmem_ld_x32_at(t_dq + 32, o + 32);
        tmem_ld_wait();
        uint8_t* dst = my_ds + row * 128;
#pragma unroll
        for (int q8 = 0; q8 < kD / 8; ++q8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * q8]) * p.scale, __uint_as_float(o[8 * q8 + 1]) * p.scale);
          v.y = pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * p.scale, __uint_as_float(o[8 * q8 + 3]) * p.scale);
          v.z = pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * p.scale, __uint_as_float(o[8 * q8 + 5]) * p.scale);
          v.w = pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * p.scale, __uint_as_float(o[8 * q8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dst + ((q8 ^ swz) << 4)) = v;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      wg_bar_sync(wg);
      if (quad == 0 && lane == 0) {
        tma_store_2d(&tmap_dq, my_ds, it.h * kD, it.b * p.T + it.q0 + wg * kTQ);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&stage_free[wg]);
      }
    }
    if (quad == 0 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

int launch_attn_bwd_dq(const AttnBwdLaunch& a, cudaStream_t stream, const char** err) {
  static thread_local char msg[192];
  *err = msg;
  msg[0] = 0;
  if (a.D != kD || a.T % kTQ != 0 || a.T <= 0) {
    snprintf(msg, sizeof(msg), "attn_bwd_dq: head_dim must be 64 and T a multiple of 128 (D=%d T=%d)",
             a.D, a.T);
    return -1;
  }
  const uint64_t rows = static_cast<uint64_t>(a.B) * a.T;
  const uint64_t width = static_cast<uint64_t>(a.H) * a.D;
  CUtensorMap tq, tdo, to, tk, tv, tdq;
  if (!make_tmap_2d(&tq, a.q, width, rows, a.ld_q, 64, kTQ) ||
      !make_tmap_2d(&tdo, a.d_o, width, rows, a.ld_do, 64, kTQ) ||
      !make_tmap_2d(&to, a.o, width, rows, a.ld_o, 64, kTQ) ||
      !make_tmap_2d(&tk, a.k, width, rows, a.ld_k, 64, kSub) ||
      !make_tmap_2d(&tv, a.v, width, rows, a.ld_v, 64, kSub) ||
      !make_tmap_2d(&tdq, a.dq, width, rows, a.ld_dq, 64, kTQ)) {
    snprintf(msg, sizeof(msg), "attn_bwd_dq: cuTensorMapEncodeTiled failed");
    return -2;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_sm100_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kTotal);
    if (e != cudaSuccess) {
      snprintf(msg, sizeof(msg), "attn_bwd_dq: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  DqParams p;
  p.B = a.B; p.T = a.T; p.H = a.H;
  p.causal = a.causal;
  p.scale = a.scale;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.n_qp = (a.T + 2 * kTQ - 1) / (2 * kTQ);
  p.n_items = a.B * a.H * p.n_qp;
  p.lse = a.lse;
  p.delta = a.delta;
  const int grid = p.n_items < num_sms ? p.n_items : num_sms;
  attn_bwd_dq_sm100_kernel<<<grid, kThreads, DqSmem::kTotal, stream>>>(tq, tdo, to, tk, tv, tdq, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(msg, sizeof(msg), "attn_bwd_dq launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace tdp
