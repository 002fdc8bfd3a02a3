// Warp-specialised persistent bf16 GEMM for sm_100a:
//   TMA (128B swizzle) -> smem ring -> tcgen05.mma (single elected thread, cta_group::1, M=128)
//   -> fp32 accumulators in TMEM (double buffered) -> tcgen05.ld epilogue with fused
//   bias / GELU / dGELU / residual / accumulate -> swizzled smem staging -> TMA store
//   (to local memory, or straight into a peer GPU's staging buffer over NVLink).
//
// This header is the shared main loop.  Plain GEMM and the GEMM+collective variants
// (all-gather->GEMM, GEMM->reduce-scatter, GEMM->all-reduce) are the same kernel with different
// tile-order / producer-wait / epilogue-destination policies selected at run time through
// GemmParams (all branches are warp uniform).
//
// Replaces the reference's torch.matmul(x, W) + in-place bias (tensor_parallel/tp_utils.py:170-174)
// and, in the fused variants, the collective that follows/precedes it (tp_utils.py:44,67,84).
#pragma once
#include "../common/ptx.cuh"

namespace tdp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;     // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
// 8 epilogue warps: two per TMEM lane quadrant (each takes one 32-column half of every 64-column
// sub-tile) -> two warps per SM sub-partition, so the epilogue math (GELU, dGELU) can hide its own
// latency and keeps up with the MMA pipe.
constexpr int kNumEpilogueWarps = 8;
constexpr int kGemmThreads = 32 * (2 + kNumEpilogueWarps);  // warp0 TMA, warp1 MMA, warps2-9 epi
// all-gather->GEMM with the push folded in: one more warp per CTA copies this CTA's share of the
// local shard to every peer while the tensor cores work on the chunks that are already there
constexpr int kGemmThreadsPush = kGemmThreads + 32;
constexpr int kMaxPeers = 8;
constexpr int kStoreCols = 64;                         // TMA-store sub-tile: 128 rows x 64 cols
constexpr int kStoreBytes = kBlockM * kStoreCols * 2;  // 16 KiB, 128B-swizzled

enum GemmAct : int {
  ACT_NONE = 0,
  ACT_GELU_TANH = 1,
  ACT_GELU_ERF = 2,
  ACT_DGELU_TANH = 3,  // out = acc * gelu'(aux_in)
  ACT_DGELU_ERF = 4,
};

enum GemmCommMode : int {
  COMM_NONE = 0,
  // A operand rows arrive chunk by chunk (all-gather): the TMA producer waits on
  // chunk_flags[chunk] >= flag_target before loading A rows of a remote chunk; the local chunk
  // is read straight from the un-gathered shard (tmap_a_local), so it never waits.
  COMM_AG_WAIT_A = 1,
  // C rows are partial sums destined for the rank that owns the row chunk: the epilogue TMA-stores
  // the tile into that rank's staging buffer (peer memory) and bumps its per-source counter.
  COMM_RS_SCATTER = 2,
};

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int a_mn_major, b_mn_major;
  // epilogue
  void* C;          // bf16 or fp32 [M, ldc]
  int ldc;
  int c_fp32;
  int accumulate;   // C += result (fp32 or bf16 read-modify-write)
  int use_tma_store;
  int epi_in_tma;    // 1: residual, 2: dGELU input -- the [128 x 64] sub-tile arrives through TMA
  int epi_aux_tma;   // pre-activation copy (aux_out) leaves through smem + TMA store
  int epi_split;     // two independent 4-warp epilogue groups, one per accumulator stage (opt-in)
  // grouped GEMM (MoE experts): output row-block m belongs to group m / grp_mblocks; the TMA
  // coordinates of the operands are shifted by group * these element offsets (0 = not grouped)
  int grp_mblocks;
  int grp_a_m, grp_a_k, grp_b_n, grp_b_k;
  int grp_bias;      // bias element offset per group
  float alpha;
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* residual;  // [M, ld_res] or null
  int ld_res;
  const __nv_bfloat16* aux_in;    // [M, ld_aux] pre-activation for dGELU
  __nv_bfloat16* aux_out;         // [M, ld_aux] pre-activation copy (value before activation)
  int ld_aux;
  int act;
  int group_m;                    // rasterisation group
  int split_k;                    // >1: K is split over CTAs, partials atomically added (fp32 C)
  int k_blocks_per_split;
  // ---- fused collective hooks ----
  int comm_mode;
  int rank, world;
  int rows_per_chunk;             // M / world (chunk c belongs to rank c)
  int has_a_local;
  uint32_t* chunk_flags;          // [world] local flags (AG wait) -- device memory of this rank
  uint32_t flag_target;           // monotonically increasing epoch value
  uint32_t* peer_tile_counter[kMaxPeers];  // per-rank counter array [world(src)]
  // in-kernel all-gather push (COMM_AG_WAIT_A, block of kGemmThreadsPush threads):
  const uint4* push_src;          // this rank's shard (push_vec 16-byte vectors), null = no push
  size_t push_vec;
  char* push_dst[kMaxPeers];      // my slot in every rank's gather buffer (incl. my own)
  char* push_mc;                  // multicast address of my slot (one store reaches all) or null
  uint32_t* push_flag[kMaxPeers]; // chunk flag [my rank] in every rank's signal pad
  uint32_t* push_ticket;          // local counter: the last CTA to finish publishes the flags
};

// C tensor maps: [0] = plain C; RS scatter: [d] = my slot in rank d's staging buffer
struct GemmStoreMaps {
  CUtensorMap m[kMaxPeers];
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
  static constexpr int kStageBytesB = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kStoreStageBytes = 2 * kStoreBytes;
  static constexpr int kBarrierBytes = 256;
  // 229 632 B: leaves > 2 KiB of the SM's 228 KiB so that a small communication CTA (push /
  // all-reduce kernels, no dynamic smem) can always be co-resident with a persistent GEMM CTA --
  // the fused all-gather->GEMM must never be able to starve the kernel it is waiting for.
  static constexpr int kTotalBytes = kStages * kStageBytes + kStoreStageBytes + kBarrierBytes;
};

// tile index -> (m_block, n_block) with grouped rasterisation so that concurrently running CTAs
// share B tiles (and a few A tiles) in L2.
TDP_DEVICE void tile_to_mn(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
  const int group = p.group_m;
  const int tiles_per_group = group * p.num_n_blocks;
  const int g = tile / tiles_per_group;
  const int first_m = g * group;
  const int rows_in_group = min(group, p.num_m_blocks - first_m);
  const int in_group = tile - g * tiles_per_group;
  m_blk = first_m + in_group % rows_in_group;
  n_blk = in_group / rows_in_group;
}

// For the collective variants tiles are ordered chunk-major: chunk order starts at a rank
// dependent offset.  AG: local chunk first, then rank+1, ... (data arrives in that order).
// RS: remote chunks first (rank+1 ...), local chunk last (it needs no transfer).
TDP_DEVICE int remap_m_block(const GemmParams& p, int m_blk) {
  if (p.comm_mode == COMM_NONE) return m_blk;
  const int blocks_per_chunk = p.rows_per_chunk / kBlockM;
  const int c = m_blk / blocks_per_chunk;
  const int r = m_blk - c * blocks_per_chunk;
  const int shift = (p.comm_mode == COMM_AG_WAIT_A) ? p.rank : (p.rank + 1);
  int chunk = c + shift;
  chunk = chunk >= p.world ? chunk - p.world : chunk;
  return chunk * blocks_per_chunk + r;
}

// Work iterator shared by the three warp roles (they must all walk the same sequence).
//   data-parallel (split_k == 1): output tiles blockIdx.x, +gridDim.x, ... over the full K range;
//   stream-K      (split_k  > 1): the flattened (tile, k-block) space is cut into equal
//   contiguous shares of `k_blocks_per_split` k-blocks per CTA; a share crosses at most a few
//   tile boundaries and every segment is accumulated into the fp32 output with vector atomics.
struct WorkIter {
  int tile, kb0, kb1;
  int pos, end;
  TDP_DEVICE explicit WorkIter(const GemmParams& p) : tile(0), kb0(0), kb1(0) {
    if (p.split_k > 1) {
      const long total = static_cast<long>(p.num_m_blocks) * p.num_n_blocks * p.num_k_blocks;
      const long b = static_cast<long>(blockIdx.x) * p.k_blocks_per_split;
      pos = static_cast<int>(b < total ? b : total);
      end = static_cast<int>(b + p.k_blocks_per_split < total ? b + p.k_blocks_per_split : total);
    } else {
      pos = blockIdx.x;
      end = p.num_m_blocks * p.num_n_blocks;
    }
  }
  TDP_DEVICE bool next(const GemmParams& p) {
    if (pos >= end) return false;
    if (p.split_k > 1) {
      tile = pos / p.num_k_blocks;
      kb0 = pos - tile * p.num_k_blocks;
      const int len = min(p.num_k_blocks - kb0, end - pos);
      kb1 = kb0 + len;
      pos += len;
    } else {
      tile = pos;
      kb0 = 0;
      kb1 = p.num_k_blocks;
      pos += gridDim.x;
    }
    return true;
  }
};

TDP_DEVICE void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Everything between the accumulator and the store for one thread-row x 32 columns, on packed
// fp32x2 lanes: v[i] holds columns (2i, 2i+1).
// `full` = the 32 columns are all inside N.
// `in_row` (optional): this thread's row of the TMA-loaded, 128B-swizzled input sub-tile (residual
// or dGELU pre-activation) -- chunk c of the 64-column sub-tile lives at (c ^ swz) * 16.
// `z_row` (optional): same layout, receives the pre-activation copy instead of a global store.
// `bias_s` (optional): the 32 bias values of these columns, staged in shared memory.
TDP_DEVICE void epilogue_math(const GemmParams& p, f32x2 (&v)[16], int row, int col0, bool full,
                              const uint8_t* in_row = nullptr, uint8_t* z_row = nullptr, int h = 0,
                              int swz = 0, const __nv_bfloat16* bias_s = nullptr) {
  if (bias_s != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 b = *reinterpret_cast<const uint4*>(bias_s + 8 * j);     // warp-wide broadcast
      v[4 * j] = add2(v[4 * j], bf16x2_to_f32x2(b.x));
      v[4 * j + 1] = add2(v[4 * j + 1], bf16x2_to_f32x2(b.y));
      v[4 * j + 2] = add2(v[4 * j + 2], bf16x2_to_f32x2(b.z));
      v[4 * j + 3] = add2(v[4 * j + 3], bf16x2_to_f32x2(b.w));
    }
  } else if (p.bias != nullptr) {
    // grouped GEMM: one bias vector per group (expert) of output rows
    const __nv_bfloat16* bias =
        p.bias + (p.grp_mblocks > 0 ? (row / (p.grp_mblocks * kBlockM)) * p.grp_bias : 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (full || col0 + 8 * j + 8 <= p.N) {
        const uint4 b = *reinterpret_cast<const uint4*>(bias + col0 + 8 * j);
        v[4 * j] = add2(v[4 * j], bf16x2_to_f32x2(b.x));
        v[4 * j + 1] = add2(v[4 * j + 1], bf16x2_to_f32x2(b.y));
        v[4 * j + 2] = add2(v[4 * j + 2], bf16x2_to_f32x2(b.z));
        v[4 * j + 3] = add2(v[4 * j + 3], bf16x2_to_f32x2(b.w));
      }
    }
  }
  if (z_row != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = f32x2_to_bf16x2(v[4 * j]); o.y = f32x2_to_bf16x2(v[4 * j + 1]);
      o.z = f32x2_to_bf16x2(v[4 * j + 2]); o.w = f32x2_to_bf16x2(v[4 * j + 3]);
      *reinterpret_cast<uint4*>(z_row + (((h * 4 + j) ^ swz) * 16)) = o;
    }
  } else if (p.aux_out != nullptr) {
    __nv_bfloat16* arow = p.aux_out + static_cast<size_t>(row) * p.ld_aux + col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (full || col0 + 8 * j + 8 <= p.N) {
        uint4 o;
        o.x = f32x2_to_bf16x2(v[4 * j]); o.y = f32x2_to_bf16x2(v[4 * j + 1]);
        o.z = f32x2_to_bf16x2(v[4 * j + 2]); o.w = f32x2_to_bf16x2(v[4 * j + 3]);
        *reinterpret_cast<uint4*>(arow + 8 * j) = o;
      }
    }
  }
  // (one branch per activation flavour: a select inside the element loop makes ptxas evaluate both)
  if (p.act == ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_tanh2(v[i]);
  } else if (p.act == ACT_GELU_ERF) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a, b;
      upk2(v[i], a, b);
      v[i] = pk2(gelu_erf(a), gelu_erf(b));
    }
  } else if (p.act == ACT_DGELU_TANH || p.act == ACT_DGELU_ERF) {
    const __nv_bfloat16* zrow = p.aux_in + static_cast<size_t>(row) * p.ld_aux + col0;
    const bool tanh_flavour = p.act == ACT_DGELU_TANH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (full || col0 + 8 * j + 8 <= p.N) {
        const uint4 z = (in_row != nullptr)
            ? *reinterpret_cast<const uint4*>(in_row + (((h * 4 + j) ^ swz) * 16))
            : *reinterpret_cast<const uint4*>(zrow + 8 * j);
        const uint32_t zw[4] = {z.x, z.y, z.z, z.w};
        if (tanh_flavour) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[4 * j + q] = dgelu_tanh2(bf16x2_to_f32x2(zw[q]), v[4 * j + q]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float z0, z1, g0, g1;
            upk2(bf16x2_to_f32x2(zw[q]), z0, z1);
            upk2(v[4 * j + q], g0, g1);
            v[4 * j + q] = pk2(g0 * dgelu_erf(z0), g1 * dgelu_erf(z1));
          }
        }
      }
    }
  }
  if (p.residual != nullptr) {
    const __nv_bfloat16* rrow = p.residual + static_cast<size_t>(row) * p.ld_res + col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (full || col0 + 8 * j + 8 <= p.N) {
        const uint4 z = (in_row != nullptr)
            ? *reinterpret_cast<const uint4*>(in_row + (((h * 4 + j) ^ swz) * 16))
            : *reinterpret_cast<const uint4*>(rrow + 8 * j);
        v[4 * j] = add2(v[4 * j], bf16x2_to_f32x2(z.x));
        v[4 * j + 1] = add2(v[4 * j + 1], bf16x2_to_f32x2(z.y));
        v[4 * j + 2] = add2(v[4 * j + 2], bf16x2_to_f32x2(z.z));
        v[4 * j + 3] = add2(v[4 * j + 3], bf16x2_to_f32x2(z.w));
      }
    }
  }
}

// accumulator registers (fp32 bit patterns from tcgen05.ld) -> packed lanes, scaled by alpha
TDP_DEVICE void epilogue_load_acc(const GemmParams& p, const uint32_t (&r)[32], f32x2 (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = pk2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  if (p.alpha != 1.f) {
    const f32x2 a2 = splat2(p.alpha);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = mul2(v[i], a2);
  }
}

// direct (row-per-thread) global store, used for fp32 output / accumulate
TDP_DEVICE void epilogue_store_direct(const GemmParams& p, const f32x2 (&v2)[16], uint8_t* c_row,
                                      int col0, bool full) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) upk2(v2[i], v[2 * i], v[2 * i + 1]);
  if (p.c_fp32) {
    float* crow = reinterpret_cast<float*>(c_row) + col0;
    if (p.split_k > 1) {
      // split-K partial: vector atomic add into the (pre-zeroed / accumulating) fp32 output
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (full || col0 + j + 4 <= p.N)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + j), "f"(v[j]),
                       "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (full || col0 + j + 4 <= p.N) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(crow + j);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(crow + j) = o;
      }
    }
  } else {
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(c_row) + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (full || col0 + j + 8 <= p.N) {
        if (p.accumulate) {
          const uint4 z = *reinterpret_cast<const uint4*>(crow + j);
          float2 t;
          t = unpack_bf16x2(z.x); v[j] += t.x; v[j + 1] += t.y;
          t = unpack_bf16x2(z.y); v[j + 2] += t.x; v[j + 3] += t.y;
          t = unpack_bf16x2(z.z); v[j + 4] += t.x; v[j + 5] += t.y;
          t = unpack_bf16x2(z.w); v[j + 6] += t.x; v[j + 7] += t.y;
        }
        uint4 o;
        o.x = pack_bf16x2(v[j], v[j + 1]); o.y = pack_bf16x2(v[j + 2], v[j + 3]);
        o.z = pack_bf16x2(v[j + 4], v[j + 5]); o.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(crow + j) = o;
      }
    }
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreadsPush, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_a_local,
                       const __grid_constant__ CUtensorMap tmap_in,
                       const __grid_constant__ CUtensorMap tmap_aux,
                       const __grid_constant__ GemmStoreMaps store_maps, const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages (<= 512)
  static_assert(kTmemCols <= 512, "TMEM overflow");

  extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B swizzle needs 1 KiB alignment
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kStageBytesA;
  uint8_t* smem_store = smem + kStages * S::kStageBytes;            // 2 x 16 KiB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + S::kStoreStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint64_t* in_bar = bars + 2 * kStages + 4;       // epilogue input sub-tile loads (2 buffers)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 6);

  const int warp_idx = threadIdx.x / 32;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;   // work items

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], p.epi_split ? kNumEpilogueWarps / 2 : kNumEpilogueWarps);
      mbar_init(&in_bar[i], 1);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<kTmemCols>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (WorkIter it(p); it.next(p);) {
        int m_blk, n_blk;
        tile_to_mn(p, it.tile, m_blk, n_blk);
        const int kb0 = it.kb0, kb1 = it.kb1;
        m_blk = remap_m_block(p, m_blk);
        int m0 = m_blk * kBlockM;
        const CUtensorMap* amap = &tmap_a;
        if (p.comm_mode == COMM_AG_WAIT_A) {
          const int chunk = m0 / p.rows_per_chunk;
          if (chunk == p.rank && p.has_a_local) {
            amap = &tmap_a_local;                 // zero-copy local shard, no wait
            m0 -= chunk * p.rows_per_chunk;
          } else {
            // wrap-safe "flag >= target" on monotonically increasing epochs
            SpinWatchdog wd;
            while (static_cast<int32_t>(ld_acquire_sys(p.chunk_flags + chunk) - p.flag_target) < 0) {
              wd.tick("all-gather chunk flag (gemm producer)", chunk, p.rank);
            }
            // peer / comm-kernel writes (generic proxy) -> our TMA reads (async proxy)
            fence_proxy_async_all();
          }
        }
        const int n0 = n_blk * BLOCK_N;
        // grouped mode: shift the operand windows to this tile's group (expert)
        const int grp = p.grp_mblocks > 0 ? m_blk / p.grp_mblocks : 0;
        const int a_m0 = m0 + grp * p.grp_a_m, a_koff = grp * p.grp_a_k;
        const int b_n0 = n0 + grp * p.grp_b_n, b_koff = grp * p.grp_b_k;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          uint8_t* sa = smem_a + stage * S::kStageBytesA;
          uint8_t* sb = smem_b + stage * S::kStageBytesB;
          const int k0 = kb * kBlockK;
          if (!p.a_mn_major) {
            tma_load_2d(amap, &full_bar[stage], sa, k0 + a_koff, a_m0);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j)
              tma_load_2d(amap, &full_bar[stage], sa + j * (64 * kBlockK * 2), a_m0 + 64 * j,
                          k0 + a_koff);
          }
          if (!p.b_mn_major) {
            tma_load_2d(&tmap_b, &full_bar[stage], sb, k0 + b_koff, b_n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (64 * kBlockK * 2), b_n0 + 64 * j,
                          k0 + b_koff);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = make_idesc_bf16_f32(kBlockM, BLOCK_N, p.a_mn_major, p.b_mn_major);
    // descriptor geometry (bytes)
    const uint32_t a_lbo = p.a_mn_major ? 64 * kBlockK * 2 : 0;
    const uint32_t b_lbo = p.b_mn_major ? 64 * kBlockK * 2 : 0;
    const uint32_t a_kstep = p.a_mn_major ? kUmmaK * 128 : kUmmaK * 2;
    const uint32_t b_kstep = p.b_mn_major ? kUmmaK * 128 : kUmmaK * 2;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (WorkIter it(p); it.next(p);) {
      // wait until the epilogue has drained this accumulator stage
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      const int kb0 = it.kb0, kb1 = it.kb1;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem_a + stage * S::kStageBytesA);
          const uint32_t sb = smem_u32(smem_b + stage * S::kStageBytesB);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = make_umma_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_umma_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
            umma_f16_ss(tmem_d, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                 // frees the smem slot when MMAs finish
          if (kb == kb1 - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator ready
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
  } else if (warp_idx < 2 + kNumEpilogueWarps && p.epi_split) {
    // ================================ epilogue, split in two groups ================================
    // Group g (warps 2-5 / 6-9, one warp per TMEM lane quadrant) drains accumulator stage g, i.e.
    // every second tile, on its own: own staging buffer, own named barrier, own TMA-store bulk
    // groups.  The two groups never wait for each other, so the exposed latencies of one (TMEM
    // load, input-tile TMA, store read-out) are filled by the other.  [opt-in: TDP_GEMM_EPI=split]
    const int g = (warp_idx - 2) >> 2;
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    const bool issuer = (quad == 2) && (lane == 0);        // warps 2 and 6 open their groups
    const int row_in_tile = quad * 32 + lane;
    const int swz = row_in_tile & 7;
    const bool in_tma = p.epi_in_tma != 0;
    const bool aux_tma = p.epi_aux_tma != 0;
    uint8_t* buf = smem_store + g * kStoreBytes;
    uint8_t* brow = buf + row_in_tile * 128;
    uint32_t in_phase = 0;
    int t = 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); };
    for (WorkIter it(p); it.next(p); ++t) {
      if ((t & 1) != g) continue;
      int m_blk, n_blk;
      tile_to_mn(p, it.tile, m_blk, n_blk);
      const int row = m_blk * kBlockM + row_in_tile;
      const int n0 = n_blk * BLOCK_N;
      const bool row_ok = row < p.M;
      const int n_sub = (min(BLOCK_N, p.N - n0) + kStoreCols - 1) / kStoreCols;
      mbar_wait(&tmem_full_bar[g], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + g * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        const int sc = sub * kStoreCols;
        if (issuer) {
          tma_store_wait_read<0>();                          // my previous store has left the buffer
          if (in_tma) {
            mbar_expect_tx(&in_bar[g], kStoreBytes);
            tma_load_2d(&tmap_in, &in_bar[g], buf, n0 + sc, m_blk * kBlockM);
          }
        }
        group_sync();
        if (in_tma) {
          mbar_wait(&in_bar[g], in_phase);
          in_phase ^= 1u;
        }
        uint32_t packed[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + sc + h * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + sc + h * 32;
          f32x2 v[16];
          epilogue_load_acc(p, r, v);
          if (row_ok && col0 < p.N)
            epilogue_math(p, v, row, col0, col0 + 32 <= p.N, in_tma ? brow : nullptr,
                          aux_tma ? brow : nullptr, h, swz);
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[h][i] = f32x2_to_bf16x2(v[i]);
        }
        if (sub == n_sub - 1) {
          tc_fence_before();                                 // last TMEM read of this tile
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
        }
        if (aux_tma) {
          // the buffer holds the pre-activation copy: send it, then reuse the buffer for C
          fence_proxy_async_smem();
          group_sync();
          if (issuer) {
            tma_store_2d(&tmap_aux, buf, n0 + sc, m_blk * kBlockM);
            tma_store_commit();
            tma_store_wait_read<0>();
          }
          group_sync();
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(brow + (((h * 4 + j) ^ swz) * 16)) =
                make_uint4(packed[h][4 * j], packed[h][4 * j + 1], packed[h][4 * j + 2],
                           packed[h][4 * j + 3]);
        fence_proxy_async_smem();
        group_sync();
        if (issuer) {
          tma_store_2d(&store_maps.m[0], buf, n0 + sc, m_blk * kBlockM);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  } else if (warp_idx < 2 + kNumEpilogueWarps) {
    // ================================ epilogue warps ================================
    // A warp may only touch TMEM lanes [32*(warp_idx%4), +32).
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;     // which 32-column half of a sub-tile this warp owns
    const int lane = threadIdx.x & 31;
    const bool issuer = (warp_idx == 2) && (lane == 0);   // the thread that owns the TMA stores
    int acc = 0;
    uint32_t acc_phase = 0;
    int store_buf = 0;
    uint32_t in_phase = 0;      // phase bits of in_bar[0..1]
    for (WorkIter it(p); it.next(p);) {
      int m_blk, n_blk;
      tile_to_mn(p, it.tile, m_blk, n_blk);
      m_blk = remap_m_block(p, m_blk);
      const int row_in_tile = quad * 32 + lane;
      const int row = m_blk * kBlockM + row_in_tile;
      const int n0 = n_blk * BLOCK_N;
      const bool row_ok = row < p.M;

      if (p.use_tma_store && p.epi_in_tma != 0 && issuer) {
        // first row-wise input sub-tile of this tile: in flight while the MMAs still run
        tma_store_wait_read<0>();                           // both staging buffers are free
        mbar_expect_tx(&in_bar[store_buf], kStoreBytes);
        tma_load_2d(&tmap_in, &in_bar[store_buf], smem_store + store_buf * kStoreBytes, n0,
                    m_blk * kBlockM);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);

      if (p.use_tma_store) {
        // ---- TMEM -> registers -> swizzled smem -> TMA store (128 rows x 64 cols at a time).
        // Row-wise epilogue operands never touch the LSU with strided 16-byte accesses: the
        // residual / dGELU input sub-tile is TMA-loaded into the staging buffer (prefetched one
        // sub-tile ahead) and combined in place; the pre-activation copy leaves through the second
        // staging buffer and its own TMA store.
        const CUtensorMap* cmap = &store_maps.m[0];
        int dst_rank = -1;
        int store_row0 = m_blk * kBlockM;
        if (p.comm_mode == COMM_RS_SCATTER) {
          dst_rank = store_row0 / p.rows_per_chunk;
          cmap = &store_maps.m[dst_rank];
          store_row0 -= dst_rank * p.rows_per_chunk;
        }
        const bool in_tma = p.epi_in_tma != 0;
        const bool aux_tma = p.epi_aux_tma != 0;
        const int swz = row_in_tile & 7;
        const int n_sub = (min(BLOCK_N, p.N - n0) + kStoreCols - 1) / kStoreCols;
#pragma unroll 1
        for (int sub = 0; sub < n_sub; ++sub) {
          const int sc = sub * kStoreCols;
          uint8_t* sbuf = smem_store + store_buf * kStoreBytes;
          uint8_t* obuf = smem_store + (store_buf ^ 1) * kStoreBytes;
          if (issuer) {
            if (in_tma) {
              tma_store_wait_read<0>();                     // store(sub-1) finished reading obuf
              if (sub + 1 < n_sub) {                        // prefetch the next input sub-tile
                mbar_expect_tx(&in_bar[store_buf ^ 1], kStoreBytes);
                tma_load_2d(&tmap_in, &in_bar[store_buf ^ 1], obuf, n0 + sc + kStoreCols,
                            m_blk * kBlockM);
              }
            } else if (aux_tma) {
              tma_store_wait_read<0>();                     // both buffers are written below
            } else {
              tma_store_wait_read<1>();                     // the store two sub-tiles ago
            }
          }
          epi_bar_sync();
          if (in_tma) {
            mbar_wait(&in_bar[store_buf], (in_phase >> store_buf) & 1u);
            in_phase ^= (1u << store_buf);
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
                            aux_tma ? srow : nullptr, h, swz);
            // row r of the staging tile: 128 bytes, 16-byte chunk c stored at (c ^ (r & 7))
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
            // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the TMA engine
          epi_bar_sync();
          if (issuer) {
            if (aux_tma) {
              tma_store_2d(&tmap_aux, sbuf, n0 + sc, m_blk * kBlockM);
              tma_store_2d(cmap, obuf, n0 + sc, store_row0);
            } else {
              tma_store_2d(cmap, sbuf, n0 + sc, store_row0);
            }
            tma_store_commit();
          }
          if (!aux_tma) store_buf ^= 1;
        }
        if (p.comm_mode == COMM_RS_SCATTER && issuer) {
          // all of this tile's bytes must have landed in the owner's memory before the counter
          // moves: wait for full completion of the bulk stores, then release at system scope.
          tma_store_wait<0>();
          fence_proxy_async_all();
          fence_acq_rel_sys();
          // unit = (32 rows x 8 columns): independent of the tile shape chosen by the host
          red_add_release_sys(p.peer_tile_counter[dst_rank] + p.rank,
                              static_cast<uint32_t>(4 * (min(BLOCK_N, p.N - n0) >> 3)));
        }
      } else {
        // ---- direct row-per-thread stores (fp32 output / accumulate)
        uint8_t* c_row = reinterpret_cast<uint8_t*>(p.C) +
                         static_cast<size_t>(row) * p.ldc * (p.c_fp32 ? 4 : 2);
#pragma unroll 1
        for (int c = half * 32; c < BLOCK_N; c += 64) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c, r);
          tmem_ld_wait();
          const int col0 = n0 + c;
          if (row_ok && col0 < p.N) {
            f32x2 v[16];
            epilogue_load_acc(p, r, v);
            const bool full = (col0 + 32 <= p.N);
            epilogue_math(p, v, row, col0, full);
            epilogue_store_direct(p, v, c_row, col0, full);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // smem must stay valid until the last bulk stores have read it
    if (issuer) tma_store_wait<0>();
  } else if (p.push_src != nullptr) {
    // ================================ all-gather push warp ================================
    // CTA b copies vectors {b*32 + lane + k*gridDim*32} of the local shard into slot `rank` of
    // every rank's gather buffer (one multimem.st, or one st per peer).  Peers' producer warps
    // wait on the chunk flag, which the last CTA to finish publishes -- the transfer overlaps the
    // MMAs on the local chunk and needs no second kernel (no co-residency assumptions).
    const int lane = threadIdx.x & 31;
    const size_t stride = static_cast<size_t>(gridDim.x) * 32;
    constexpr int kUnroll = 8;
    for (size_t i = static_cast<size_t>(blockIdx.x) * 32 + lane; i < p.push_vec;
         i += stride * kUnroll) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t idx = i + u * stride;
        if (idx < p.push_vec) v[u] = ld_nc_v4(p.push_src + idx);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t idx = i + u * stride;
        if (idx < p.push_vec) {
          if (p.push_mc != nullptr) {
            multimem_st_v4(p.push_mc + idx * 16, v[u]);
          } else {
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
              if (r < p.world) st_na_v4(p.push_dst[r] + idx * 16, v[u]);
          }
        }
      }
    }
    fence_acq_rel_sys();                         // every lane drains its own multicast stores
    __syncwarp();
    uint32_t last = 0;
    if (lane == 0) {
      fence_acq_rel_sys();                       // my stores happen-before the ticket
      const uint32_t t = atomicAdd(p.push_ticket, 1u);
      fence_acq_rel_sys();                       // the other CTAs' stores happen-before the flags
      last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      if (lane < p.world) st_release_sys(p.push_flag[lane], p.flag_target);
      if (lane == 0) *p.push_ticket = 0u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace tdp
