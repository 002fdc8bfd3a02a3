// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on one TPC
// computes a 256 x {256|128} output tile.  Each CTA TMA-loads its own 128 rows of A and one 128-column
// half of B per k-block (32 KiB / stage instead of 48 KiB -> 6 stages, and every B byte is read
// from L2 once per pair instead of once per CTA); the leader CTA's MMA warp issues
// `tcgen05.mma.cta_group::2` (M = 256) which reads both CTAs' shared memory and writes both CTAs'
// TMEM; `tcgen05.commit ... multicast::cluster` releases the smem slots / publishes the
// accumulators in both CTAs; each CTA runs the epilogue for its own 128 rows.
//
// Plain GEMMs only (no collective mode, no stream-K, bf16 TMA-store epilogue); everything else
// uses the 1-CTA kernel in gemm_sm100.cuh, whose helpers are shared.
#pragma once
#include "gemm_sm100.cuh"

namespace tdp {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address -> CTA 0 of the pair

TDP_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
TDP_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is reported to the *leader* CTA's mbarrier
TDP_DEVICE void tma_load_2d_2sm(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0,
                                int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
TDP_DEVICE void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once) on the same barrier offset in both CTAs of the pair when the MMAs have completed
TDP_DEVICE void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the leader CTA's copy of a barrier (works from either CTA)
TDP_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  // relaxed: what is handed over is a TMEM accumulator stage, already ordered by
  // tcgen05.wait::ld + tcgen05.fence::before_thread_sync -- a cluster-scope *release* would also
  // drain this warp's outstanding shared-memory stores (an ERRBAR per warp and tile)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                   smem_u32(bar) & kPeerBitMask)
               : "memory");
}
template <uint32_t kCols>
TDP_DEVICE void tmem_alloc_2sm(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_holder)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
TDP_DEVICE void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}

// kEpiBufs = number of 16 KiB epilogue staging buffers.  2: the main loop gets the shared memory
// (6 / 8 operand stages) -- right for long-K products.  6 (256-wide tiles only): 4 operand stages
// and a six-deep staging ring, for short-K products with a heavy fused epilogue (GELU + second
// output, GELU' x gradient, residual), where the chain "row-wise input TMA load -> math -> TMA
// store" of one 64-column sub-tile is longer than the sub-tile's share of the MMA time: with two
// buffers every sub-tile waits for the previous store to be read out and for its own input to
// land; with six, inputs are prefetched three sub-tiles ahead (across tile boundaries) and up to
// three stores are in flight.
template <int BLOCK_N_, int kEpiBufs_ = 2>
struct Gemm2CtaSmem {
  static constexpr int kBlockN = BLOCK_N_;                         // per CTA pair: 256 or 128
  static constexpr int kStoreBufs = kEpiBufs_;
  static constexpr int kStageBytesA = kBlockM * kBlockK * 2;       // 16 KiB (my 128 rows)
  static constexpr int kStageBytesB = (kBlockN / 2) * kBlockK * 2; // 16 / 8 KiB (my half of N)
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = kBlockN == 256 ? (kEpiBufs_ == 2 ? 6 : 4) : 8;
  static constexpr int kStoreStageBytes = kStoreBufs * kStoreBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kBiasBytes = 2 * kBlockN * 2;      // bias of the current / next tile's columns
  static constexpr int kTotalBytes =
      kStages * kStageBytes + kStoreStageBytes + kBarrierBytes + kBiasBytes;
  static_assert(kEpiBufs_ == 2 || kEpiBufs_ == 6, "staging ring depth");
};

// p.num_m_blocks counts 256-row blocks here; p.num_n_blocks BLOCK_N-column blocks.
// BLOCK_N = 128 gives 256 x 128 pair tiles: twice as many tiles for weight-gradient shaped
// products (few output tiles, long K) that would otherwise leave half of the machine idle.
template <int BLOCK_N, int kEpiBufs = 2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a,
                            const __grid_constant__ CUtensorMap tmap_b,
                            const __grid_constant__ CUtensorMap tmap_in,
                            const __grid_constant__ CUtensorMap tmap_aux,
                            const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using S = Gemm2CtaSmem<BLOCK_N, kEpiBufs>;
  constexpr int kStages = S::kStages;
  constexpr int NB = S::kStoreBufs;          // staging ring depth
  constexpr int kPrefetch = NB / 2;          // row-wise inputs are loaded this many sub-tiles ahead
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kStageBytesA;
  uint8_t* smem_store = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + S::kStoreStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint64_t* in_bar = bars + 2 * kStages + 4;                     // NB
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + NB);
  __nv_bfloat16* smem_bias =
      reinterpret_cast<__nv_bfloat16*>(smem_store + S::kStoreStageBytes + S::kBarrierBytes);

  const int warp_idx = threadIdx.x / 32;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);      // leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[i], 1);     // multicast tcgen05.commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * kNumEpilogueWarps);   // epilogue warps of both CTAs
    }
    for (int i = 0; i < NB; ++i) mbar_init(&in_bar[i], 1);
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc_2sm<kTmemCols>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_to_mn(p, tile, m_blk, n_blk);
        const int m0 = m_blk * 2 * kBlockM + static_cast<int>(cta_rank) * kBlockM;
        const int n0 = n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / 2);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
          uint8_t* sa = smem_a + stage * S::kStageBytesA;
          uint8_t* sb = smem_b + stage * S::kStageBytesB;
          const int k0 = kb * kBlockK;
          if (!p.a_mn_major) {
            tma_load_2d_2sm(&tmap_a, &full_bar[stage], sa, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j)
              tma_load_2d_2sm(&tmap_a, &full_bar[stage], sa + j * (64 * kBlockK * 2), m0 + 64 * j, k0);
          }
          if (!p.b_mn_major) {
            tma_load_2d_2sm(&tmap_b, &full_bar[stage], sb, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < (BLOCK_N / 2) / 64; ++j)
              tma_load_2d_2sm(&tmap_b, &full_bar[stage], sb + j * (64 * kBlockK * 2), n0 + 64 * j, k0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (leader) {
      const uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, BLOCK_N, p.a_mn_major, p.b_mn_major);
      const uint32_t a_lbo = p.a_mn_major ? 64 * kBlockK * 2 : 0;
      const uint32_t b_lbo = p.b_mn_major ? 64 * kBlockK * 2 : 0;
      const uint32_t a_kstep = p.a_mn_major ? kUmmaK * 128 : kUmmaK * 2;
      const uint32_t b_kstep = p.b_mn_major ? kUmmaK * 128 : kUmmaK * 2;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem_a + stage * S::kStageBytesA);
            const uint32_t sb = smem_u32(smem_b + stage * S::kStageBytesB);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = make_umma_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
              const uint64_t db = make_umma_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
              umma_f16_ss_2sm(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == p.num_k_blocks - 1) umma_commit_2sm(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp_idx < 2 + kNumEpilogueWarps) {
    // ================================ epilogue warps (both CTAs, own 128 rows) ================
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    const int lane = threadIdx.x & 31;
    const bool issuer = (warp_idx == 2) && (lane == 0);
    const bool in_tma = p.epi_in_tma != 0;
    const bool aux_tma = p.epi_aux_tma != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // The staging buffers form a ring indexed by the running sub-tile number q (it keeps counting
    // across tiles).  plain: buffer q % NB, a store may be read out while NB-1 newer ones are
    // queued.  aux (two outputs per sub-tile): buffer pairs, NB/2-1 newer groups.  in (row-wise
    // input): input(q + kPrefetch) is loaded into buffer (q + kPrefetch) % NB as soon as
    // store(q + kPrefetch - NB) has been read out, i.e. with NB-kPrefetch-1 newer groups pending.
    uint32_t q = 0;
    uint32_t in_phase = 0;                       // phase bit per in_bar
    // issuer only: prefetch cursor, kPrefetch sub-tiles ahead of the compute cursor
    int pf_tile = cluster_id, pf_sub = 0;
    uint32_t pf_q = 0;
    auto pf_issue = [&]() {
      if (pf_tile >= num_tiles) return;
      int pm, pn;
      tile_to_mn(p, pf_tile, pm, pn);
      const int pm0 = pm * 2 * kBlockM + static_cast<int>(cta_rank) * kBlockM;
      const int pn0 = pn * BLOCK_N;
      const int pns = (min(BLOCK_N, p.N - pn0) + kStoreCols - 1) / kStoreCols;
      const int b = static_cast<int>(pf_q % NB);
      mbar_expect_tx(&in_bar[b], kStoreBytes);
      tma_load_2d(&tmap_in, &in_bar[b], smem_store + b * kStoreBytes, pn0 + pf_sub * kStoreCols, pm0);
      ++pf_q;
      if (++pf_sub == pns) {
        pf_sub = 0;
        pf_tile += num_clusters;
      }
    };
    if (in_tma && issuer) {
#pragma unroll 1
      for (int i = 0; i < kPrefetch; ++i) pf_issue();      // the ring is empty: no wait needed
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_to_mn(p, tile, m_blk, n_blk);
      const int m0 = m_blk * 2 * kBlockM + static_cast<int>(cta_rank) * kBlockM;
      const int row_in_tile = quad * 32 + lane;
      const int row = m0 + row_in_tile;
      const int n0 = n_blk * BLOCK_N;
      const bool row_ok = row < p.M;
      const int swz = row_in_tile & 7;

      // bias of this tile's columns -> smem while the MMAs still run (a global load per sub-tile
      // in every thread was the top stall of the fused-GELU epilogue); double buffered by tile
      // parity, ordered by the sub-tile barrier below
      __nv_bfloat16* bias_s = nullptr;
      if (p.bias != nullptr && p.grp_mblocks == 0) {
        bias_s = smem_bias + acc * BLOCK_N;
        const int t = threadIdx.x - 64;                        // 0 .. 255 over the epilogue warps
        if (t < BLOCK_N) bias_s[t] = (n0 + t < p.N) ? p.bias[n0 + t] : __float2bfloat16(0.f);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
      const int n_sub = (min(BLOCK_N, p.N - n0) + kStoreCols - 1) / kStoreCols;
      if (bias_s != nullptr && (NB >= 4)) epi_bar_sync();      // (NB == 2: the first barrier below)
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub, ++q) {
        const int sc = sub * kStoreCols;
        const int b0 = aux_tma ? static_cast<int>(2 * (q % (NB / 2))) : static_cast<int>(q % NB);
        uint8_t* sbuf = smem_store + b0 * kStoreBytes;
        uint8_t* obuf = smem_store + (b0 ^ 1) * kStoreBytes;      // second output (aux mode only)
        if (issuer && in_tma) {
          tma_store_wait_read<NB - kPrefetch - 1>();
          pf_issue();                                         // input of sub-tile q + kPrefetch
        }
        if constexpr (NB < 4) {
          // two buffers: everybody waits until the issuer has seen the buffer read out
          if (issuer && !in_tma) {
            if (aux_tma) tma_store_wait_read<NB / 2 - 1>();
            else tma_store_wait_read<NB - 1>();
          }
          epi_bar_sync();
        }
        // (deep ring: the buffer of sub-tile q was confirmed free by the issuer before the closing
        //  barrier of sub-tile q-1, see below -- one barrier per sub-tile; in input mode the
        //  buffer's own mbarrier is the hand-over)
        if (in_tma) {
          mbar_wait(&in_bar[b0], (in_phase >> b0) & 1u);
          in_phase ^= (1u << b0);
        }
        uint8_t* srow = sbuf + row_in_tile * 128;
        uint8_t* orow = obuf + row_in_tile * 128;
        {
          const int h = half;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + sc + h * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + sc + h * 32;
          f32x2 v[16];
          epilogue_load_acc(p, r, v);
          if (row_ok && col0 < p.N)
            epilogue_math(p, v, row, col0, col0 + 32 <= p.N, in_tma ? srow : nullptr,
                          aux_tma ? srow : nullptr, h, swz,
                          bias_s != nullptr ? bias_s + sc + h * 32 : nullptr);
          uint8_t* wrow = aux_tma ? orow : srow;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = f32x2_to_bf16x2(v[4 * j]);
            o.y = f32x2_to_bf16x2(v[4 * j + 1]);
            o.z = f32x2_to_bf16x2(v[4 * j + 2]);
            o.w = f32x2_to_bf16x2(v[4 * j + 3]);
            *reinterpret_cast<uint4*>(wrow + (((h * 4 + j) ^ swz) * 16)) = o;
          }
        }
        if (sub == n_sub - 1) {
          // accumulator drained in this CTA: tell the leader's MMA warp (it needs both CTAs)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
        }
        fence_proxy_async_smem();
        if constexpr (NB >= 4) {
          // the buffer (pair) of the NEXT sub-tile must have been read out: NB-2 (NB/2-2) newer
          // store groups may still be pending at this point (store(q) is committed below)
          if (issuer && !in_tma) {
            if (aux_tma) tma_store_wait_read<NB / 2 - 2>();
            else tma_store_wait_read<NB - 2>();
          }
        }
        epi_bar_sync();
        if (issuer) {
          if (aux_tma) {
            tma_store_2d(&tmap_aux, sbuf, n0 + sc, m0);
            tma_store_2d(&tmap_c, obuf, n0 + sc, m0);
          } else {
            tma_store_2d(&tmap_c, sbuf, n0 + sc, m0);
          }
          tma_store_commit();
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still be reading my smem / signalling me
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

}  // namespace tdp
