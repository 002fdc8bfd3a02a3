// Flash-attention forward for sm_100a: persistent, warp-specialised, tcgen05.mma with S *and* the
// running output O in TMEM, TMA-fed K / V ring, two softmax warp-groups (one 128-row query tile
// each) ping-ponging on the tensor core, lazy (thresholded) online-softmax rescaling.
//
//   warps 0-3    softmax warp-group 0: query rows q0 .. q0+127      (thread = row = TMEM lane)
//   warps 4-7    softmax warp-group 1: query rows q0+128 .. q0+255
//   warp  8      TMA producer  (Q tiles per work item; K(j), V(j) through a 3-stage ring each)
//   warp  9      MMA issuer    (one elected thread) + TMEM allocation
//   warps 10-11  idle (they only exist so that warps 8-11 form a warp-group for setmaxnreg:
//                the two softmax groups run with 216 registers, the rest with 72: (168-72)*128 = (216-168)*256 registers change hands)
//
//   work item = (batch b, head h, 256 query rows); one CTA per SM walks a static, heavy-first
//   "snake" schedule over all items (causal tiles differ 5x in work), so set-up cost (TMEM
//   allocation, barrier init, tensor-map prefetch) is paid once per SM, not once per tile, and the
//   TMA / MMA warps run ahead into the next item while the softmax groups finish the current one.
//
//   per KV tile j and warp-group w:
//     S_w  = Q_w K_j^T           UMMA 128x128x64   -> TMEM cols [128w, 128w+128)
//     WG w: one TMEM read of the whole 128-column row (the TMEM region is released at once, so
//           S_w(j+1) is computed while the group still works on tile j), row max; the running max m is only
//           advanced (and O_w / l rescaled, in TMEM) when the new max exceeds it by > 2^8 --
//           rare after the first tile, so the O correction is off the critical path
//           P = exp2(S*c - m) -> bf16 -> swizzled smem (A operand)
//     O_w += P_w V_j             UMMA 128x64x128   -> TMEM cols [256+64w, ...)  (accumulated)
//   end of item: O_w / l -> bf16 -> swizzled smem -> TMA store; LSE (natural log) -> global.
//
// Layout contract: q, k, v, o are [B*T, ld] row-major "token matrices" whose row r = b*T + t holds
// all heads of a token (head h at columns col0 + h*64 ...): exactly the packed qkv GEMM output
// (q | k | v along the row) and the [B, T, H*D] attention output -- no permutes, no split copies.
// head_dim = 64, T % 128 == 0.
//
// Validated on B200 by scripts/attn_check.py and tests/test_gpu_kernels.py (vs an fp32 dense
// reference).  (reference: parallel/tensor_parallel/attn.py:40-43 is the unfused QK^T / softmax /
// PV this replaces; explore/flash-attn/tile_attn.py:100-212 is the tiled algorithm.)
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

constexpr int kHeadDim = 64;
constexpr int kTileQ = 128;          // query rows per warp-group
constexpr int kTileKV = 128;         // keys per iteration
constexpr int kStagesKV = 3;
constexpr int kAttnThreads = 32 * 12;
constexpr int kTileBytes = kTileQ * kHeadDim * 2;        // 16 KiB: one [128 x 64] bf16 tile
constexpr int kPBytes = kTileQ * kTileKV * 2;            // 32 KiB: P tile, two K-major k-blocks
constexpr uint32_t kTmemColsAttn = 512;                  // S0 | S1 | O0 | O1 (384 used)
constexpr float kRescaleThreshold = 8.f;                 // log2 units

struct AttnSmem {
  static constexpr int kQ = 0;                                   // 2 buffers x 2 tiles
  static constexpr int kK = kQ + 4 * kTileBytes;                 // kStagesKV tiles
  static constexpr int kV = kK + kStagesKV * kTileBytes;         // kStagesKV tiles
  static constexpr int kP = kV + kStagesKV * kTileBytes;         // 2 P tiles
  static constexpr int kBars = kP + 2 * kPBytes;
  static constexpr int kTotal = kBars + 256;                     // 229 632 B
};

struct AttnParams {
  int B, T, H;
  int causal;
  float scale_log2;        // softmax scale * log2(e)
  int n_qp;                // 256-row query blocks per (b, h)
  int n_items;             // B * H * n_qp
  float* lse;              // [B, H, T]
};

using namespace attn;

// Static schedule shared by all roles: items are ordered heavy-first (for causal attention the
// last query block of every (b, h) first), CTA c takes position c of even rounds and position
// G-1-c of odd rounds ("snake"), which evens out the 5x spread in item weight.
struct Item {
  int b, h, q0;
  int n_kv[2];
  int n_max;
};
TDP_DEVICE bool get_item(const AttnParams& p, int round, Item& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = round * G + ((round & 1) ? (G - 1 - c) : c);
  if (idx >= p.n_items) return false;
  const int bh_count = p.B * p.H;
  const int qp = p.n_qp - 1 - idx / bh_count;
  const int bh = idx - (idx / bh_count) * bh_count;
  it.b = bh / p.H;
  it.h = bh - it.b * p.H;
  it.q0 = qp * 2 * kTileQ;
  const int n_all = p.T / kTileKV;
  it.n_kv[0] = p.causal ? (it.q0 / kTileKV + 1) : n_all;
  it.n_kv[1] = (it.q0 + kTileQ < p.T) ? (p.causal ? (it.q0 / kTileKV + 2) : n_all) : 0;
  it.n_max = it.n_kv[0] > it.n_kv[1] ? it.n_kv[0] : it.n_kv[1];
  return true;
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
  uint64_t* q_full = bars;                       // 2 (Q tiles of item `round` live in buffer round & 1)
  uint64_t* q_empty = bars + 2;                  // 2: every S MMA of the item has completed
  uint64_t* k_full = bars + 4;                   // kStagesKV
  uint64_t* k_empty = k_full + kStagesKV;
  uint64_t* v_full = k_empty + kStagesKV;
  uint64_t* v_empty = v_full + kStagesKV;
  uint64_t* s_full = v_empty + kStagesKV;        // 2: S_w ready for warp-group w
  uint64_t* s_free = s_full + 2;                 // 2: S_w is in registers, TMEM region reusable
  uint64_t* p_ready = s_free + 2;                // 2: P_w written (O_w consistent)
  uint64_t* o_full = p_ready + 2;                // 2: O_w += P_w V complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 8 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < kStagesKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&s_free[w], 4);       // one elected lane per warp of the warp-group
      mbar_init(&p_ready[w], 4);
      mbar_init(&o_full[w], 1);
    }
    fence_barrier_init();
  } else if (warp_idx == 9) {
    tmem_alloc<kTmemColsAttn>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp_idx == 8) {
      // ================================ TMA producer ================================
      if (elect_one()) {
        uint32_t g = 0;                    // K/V ring position, runs across items
        Item it;
        for (int round = 0; get_item(p, round, it); ++round) {
          const int row_base = it.b * p.T;
          // Q is double buffered: the next item's query tiles load while this item computes
          const int qb = round & 1;
          if (round >= 2) mbar_wait(&q_empty[qb], ((round >> 1) - 1) & 1);
          const int n_q = it.n_kv[1] > 0 ? 2 : 1;
          mbar_expect_tx(&q_full[qb], n_q * kTileBytes);
          for (int w = 0; w < n_q; ++w)
            tma_load_2d(&tmap_q, &q_full[qb], smem_q + (2 * qb + w) * kTileBytes, it.h * kHeadDim,
                        row_base + it.q0 + w * kTileQ);
          for (int j = 0; j < it.n_max; ++j, ++g) {
            const int st = g % kStagesKV;
            const uint32_t ph = (g / kStagesKV) & 1;
            mbar_wait(&k_empty[st], ph ^ 1);
            mbar_expect_tx(&k_full[st], kTileBytes);
            tma_load_2d(&tmap_k, &k_full[st], smem_k + st * kTileBytes, it.h * kHeadDim,
                        row_base + j * kTileKV);
            mbar_wait(&v_empty[st], ph ^ 1);
            mbar_expect_tx(&v_full[st], kTileBytes);
            // one [128 keys x 64 d] box: rows of 128 B, i.e. two stacked [64 x 64] boxes -- exactly
            // the MN-major B operand of P.V (one 8 KiB box per 64-key k-block)
            tma_load_2d(&tmap_v, &v_full[st], smem_v + st * kTileBytes, it.h * kHeadDim,
                        row_base + j * kTileKV);
          }
        }
      }
    } else if (warp_idx == 9) {
      // ================================ MMA issuer ================================
      // S = Q K^T : A = Q  [128 x 64]  K-major,  B = K [128 keys x 64] K-major   (N = 128)
      // O+= P V   : A = P  [128 x 128] K-major (2 k-blocks), B = V [128 keys x 64] MN-major (N = 64)
      const uint32_t idesc_s = make_idesc_bf16_f32(kTileQ, kTileKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16_f32(kTileQ, kHeadDim, 0, 1);
      constexpr uint32_t kUmmaKBytes = 16 * 2;          // K = 16 bf16 along a 128-byte swizzled row
      int qbuf = 0;
      auto issue_s = [&](int w, int st) {
        const uint32_t sa = smem_u32(smem_q + (2 * qbuf + w) * kTileBytes);
        const uint32_t sb = smem_u32(smem_k + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          const uint64_t da = make_umma_smem_desc_sw128(sa + k * kUmmaKBytes, 0, 1024);
          const uint64_t db = make_umma_smem_desc_sw128(sb + k * kUmmaKBytes, 0, 1024);
          umma_f16_ss(tmem_base + w * kTileKV, da, db, idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[w]);
      };
      auto issue_pv = [&](int w, int st, bool accumulate) {
        const uint32_t sp = smem_u32(smem_p + w * kPBytes);
        const uint32_t sv = smem_u32(smem_v + st * kTileBytes);
#pragma unroll
        for (int kb = 0; kb < kTileKV / 64; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A: k-block kb of P is a [128 x 64] K-major tile; B: box kb of V, 16 key rows per step
            const uint64_t da =
                make_umma_smem_desc_sw128(sp + kb * kTileBytes + k * kUmmaKBytes, 0, 1024);
            const uint64_t db = make_umma_smem_desc_sw128(
                sv + kb * (kTileBytes / 2) + k * 16 * 128, 64 * 64 * 2, 1024);
            umma_f16_ss(tmem_base + 2 * kTileKV + w * kHeadDim, da, db, idesc_o,
                        (accumulate || (kb | k) != 0) ? 1u : 0u);
          }
        }
        umma_commit(&o_full[w]);
      };

      // Event-driven issue loop (one thread): per warp-group two event streams,
      //   s_free[w]  (S_w(j) has been read into registers)  -> issue S_w(j+1)
      //   p_ready[w] (P_w(j) is in smem)                     -> issue O_w += P_w(j) V(j)
      // served in arrival order by polling, so the two softmax groups drift half a period apart
      // (true ping-pong) and S_w(j+1) is computed while group w still works on tile j.
      if (elect_one()) {
        uint32_t g = 0;
        uint32_t a_cnt[2] = {0u, 0u}, b_cnt[2] = {0u, 0u};     // events consumed (phase parity)
        Item it;
        for (int round = 0; get_item(p, round, it); ++round) {
          int a_loc[2] = {0, 0}, b_loc[2] = {0, 0};            // per item
          int s_iss[2] = {0, 0};                               // S_w(0..s_iss-1) issued
          int rk = 0, rv = 0;                                  // K / V tiles released so far
          bool q_released = false;
          auto release = [&]() {
            // K(jj) is free once every group that needs it has had S(jj) issued, V(jj) once
            // every group has had its P.V(jj) issued (commit = when those MMAs complete)
            while (rk < it.n_max && (s_iss[0] > rk || rk >= it.n_kv[0]) &&
                   (s_iss[1] > rk || rk >= it.n_kv[1])) {
              umma_commit(&k_empty[(g + rk) % kStagesKV]);
              ++rk;
            }
            while (rv < it.n_max && (b_loc[0] > rv || rv >= it.n_kv[0]) &&
                   (b_loc[1] > rv || rv >= it.n_kv[1])) {
              umma_commit(&v_empty[(g + rv) % kStagesKV]);
              ++rv;
            }
            if (!q_released && s_iss[0] >= it.n_kv[0] && s_iss[1] >= it.n_kv[1]) {
              umma_commit(&q_empty[qbuf]);                     // last S of the item issued
              q_released = true;
            }
          };
          qbuf = round & 1;
          mbar_wait(&q_full[qbuf], (round >> 1) & 1);
          mbar_wait(&k_full[g % kStagesKV], (g / kStagesKV) & 1);
          tc_fence_after();
          for (int w = 0; w < 2; ++w)
            if (it.n_kv[w] > 0) { issue_s(w, g % kStagesKV); s_iss[w] = 1; }
          release();
          while (a_loc[0] < it.n_kv[0] || b_loc[0] < it.n_kv[0] || a_loc[1] < it.n_kv[1] ||
                 b_loc[1] < it.n_kv[1]) {
            for (int w = 0; w < 2; ++w) {
              // an event is only consumed when the K / V stage it needs has landed: a blocking
              // wait here could starve the other group, whose progress frees ring stages
              if (a_loc[w] < it.n_kv[w] && mbar_test_wait(&s_free[w], a_cnt[w] & 1)) {
                const int j = a_loc[w];
                const uint32_t gj = g + j + 1;
                if (j + 1 >= it.n_kv[w]) {
                  ++a_cnt[w];
                  ++a_loc[w];
                } else if (mbar_test_wait(&k_full[gj % kStagesKV], (gj / kStagesKV) & 1)) {
                  ++a_cnt[w];
                  ++a_loc[w];
                  tc_fence_after();
                  issue_s(w, gj % kStagesKV);
                  s_iss[w] = j + 2;
                  release();
                }
              }
              if (b_loc[w] < it.n_kv[w] && mbar_test_wait(&p_ready[w], b_cnt[w] & 1)) {
                const int j = b_loc[w];
                const uint32_t gj = g + j;
                if (mbar_test_wait(&v_full[gj % kStagesKV], (gj / kStagesKV) & 1)) {
                  ++b_cnt[w];
                  tc_fence_after();
                  issue_pv(w, gj % kStagesKV, j > 0);
                  b_loc[w] = j + 1;
                  release();
                }
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int wg = warp_idx >> 2;
    const int quad = warp_idx & 3;
    const int row = quad * 32 + lane;                       // row in the 128-row tile = TMEM lane
    const uint32_t t_s = tmem_base + wg * kTileKV + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o =
        tmem_base + 2 * kTileKV + wg * kHeadDim + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* my_p = smem_p + wg * kPBytes;
    const int swz = row & 7;
    uint32_t cnt_s = 0, cnt_o = 0;
    bool store_pending = false;                             // my_p is being read by a TMA store
    Item it;
    for (int round = 0; get_item(p, round, it); ++round) {
      const int n_mine = it.n_kv[wg];
      if (n_mine == 0) continue;
      float m_run = 0.f, l_run = 0.f;
      for (int j = 0; j < n_mine; ++j) {
        mbar_wait(&s_full[wg], cnt_s & 1);
        ++cnt_s;
        tc_fence_after();
        uint32_t r[kTileKV];
#pragma unroll
        for (int c = 0; c < kTileKV / 32; ++c) tmem_ld_x32_at(t_s + c * 32, r + c * 32);
        tmem_ld_wait();
        // S_w lives in registers now: release the TMEM region so that S_w(j+1) is computed while
        // this group is still busy with tile j (the registers are the second buffer)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);
        bool pv_done = (j == 0);            // has P.V(j-1) been waited for (P tile free, O final)?
        if (p.causal && j == n_mine - 1) {                  // keys j*128 + c vs query q0+128wg+row
#pragma unroll
          for (int i = 0; i < kTileKV; ++i)
            if (i > row) r[i] = 0xff800000u;                // -inf
        }
        // 8 independent 3-input max chains (a single chain would be 63 dependent ops)
        float mx[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          mx[c] = fmax3(__uint_as_float(r[3 * c]), __uint_as_float(r[3 * c + 1]),
                        __uint_as_float(r[3 * c + 2]));
#pragma unroll
        for (int i = 24; i + 15 < kTileKV; i += 16) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            mx[c] = fmax3(mx[c], __uint_as_float(r[i + 2 * c]), __uint_as_float(r[i + 2 * c + 1]));
        }
        // 24 + 6*16 = 120: the last 8 columns
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], __uint_as_float(r[120 + c]));
        const float sm = fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]),
                               fmaxf(mx[6], mx[7])) * p.scale_log2;
        if (j == 0) {
          m_run = sm;
        } else if (__any_sync(0xffffffffu, sm > m_run + kRescaleThreshold)) {
          // rare: advance the running max and rescale O (in TMEM) and l
          mbar_wait(&o_full[wg], cnt_o & 1);      // O_w must hold all of P.V(0..j-1)
          ++cnt_o;
          pv_done = true;
          tc_fence_after();
          const float m_new = fmaxf(m_run, sm);
          const float alpha = ex2(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
#pragma unroll
          for (int c = 0; c < kHeadDim / 32; ++c) {
            uint32_t o[32];
            tmem_ld_x32_at(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(t_o + c * 32, o);
          }
          tmem_st_wait();
        }
        // ---- P = exp2(S * c - m) (in registers, as bf16 pairs), row sum
        float ls0 = 0.f, ls1 = 0.f;
        const float neg_m = -m_run;
        uint32_t pk[kTileKV / 2];
#pragma unroll
        for (int i = 0; i < kTileKV; i += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, neg_m));
          const float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, neg_m));
          ls0 += p0;
          ls1 += p1;
          pk[i / 2] = pack_bf16x2(p0, p1);
        }
        l_run += ls0 + ls1;
        // ---- the P tile is free once P.V(j-1) has completed (issued a whole softmax ago)
        if (!pv_done) {
          mbar_wait(&o_full[wg], cnt_o & 1);
          ++cnt_o;
        }
        // ... and once the O tile of the previous item has left through it
        if (store_pending) {
          if (quad == 0 && lane == 0) tma_store_wait_read<0>();
          wg_bar_sync(wg);
          store_pending = false;
        }
        // bf16 P into the K-major swizzled A tile: columns [32c, 32c+32) = k-block c/2,
        // 16-byte chunks (c%2)*4 .. +3 of this row
#pragma unroll
        for (int c = 0; c < kTileKV / 32; ++c) {
          uint8_t* dst = my_p + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (c & 1) * 4 + q4;
            *reinterpret_cast<uint4*>(dst + ((chunk ^ swz) << 4)) =
                make_uint4(pk[16 * c + 4 * q4], pk[16 * c + 4 * q4 + 1], pk[16 * c + 4 * q4 + 2],
                           pk[16 * c + 4 * q4 + 3]);
          }
        }
        // O_w is consistent, P_w is written: hand them to the MMA warp
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[wg]);
      }

      // ---- end of item: O / l -> bf16 -> swizzled staging (my_p: the last P.V has completed) -> TMA
      mbar_wait(&o_full[wg], cnt_o & 1);
      ++cnt_o;
      tc_fence_after();
      {
        uint32_t o[kHeadDim];
#pragma unroll
        for (int c = 0; c < kHeadDim / 32; ++c) tmem_ld_x32_at(t_o + c * 32, o + c * 32);
        tmem_ld_wait();
        const float inv_l = 1.f / l_run;
        uint8_t* dst = my_p + row * 128;
#pragma unroll
        for (int q8 = 0; q8 < kHeadDim / 8; ++q8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * q8]) * inv_l, __uint_as_float(o[8 * q8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * inv_l, __uint_as_float(o[8 * q8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * inv_l, __uint_as_float(o[8 * q8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * inv_l, __uint_as_float(o[8 * q8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + ((q8 ^ swz) << 4)) = v;
        }
      }
      const int t = it.q0 + wg * kTileQ + row;
      if (p.lse != nullptr)
        p.lse[(static_cast<size_t>(it.b) * p.H + it.h) * p.T + t] =
            (m_run + lg2(l_run)) * 0.6931471805599453f;
      tc_fence_before();                  // O_w has been read: the next item's first P.V may overwrite
      fence_proxy_async_smem();
      wg_bar_sync(wg);
      if (quad == 0 && lane == 0) {
        tma_store_2d(&tmap_o, my_p, it.h * kHeadDim, it.b * p.T + it.q0 + wg * kTileQ);
        tma_store_commit();
      }
      store_pending = true;
    }
    if (quad == 0 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) {
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
  static int num_sms = 0;
  if (num_sms == 0) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) {
      snprintf(msg, sizeof(msg), "attn_fwd: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  AttnParams p;
  p.B = a.B; p.T = a.T; p.H = a.H;
  p.causal = a.causal;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.n_qp = (a.T + 2 * kTileQ - 1) / (2 * kTileQ);
  p.n_items = a.B * a.H * p.n_qp;
  p.lse = a.lse;
  const int grid = p.n_items < num_sms ? p.n_items : num_sms;
  attn_fwd_sm100_kernel<<<grid, kAttnThreads, AttnSmem::kTotal, stream>>>(tq, tk, tv, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(msg, sizeof(msg), "attn_fwd launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace tdp
