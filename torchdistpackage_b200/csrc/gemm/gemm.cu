// Host side of the sm_100a tcgen05 GEMM: TMA tensor-map construction (cached), tile-shape
// selection, persistent launch.  Kernel body: gemm_sm100.cuh.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>

#include "../common/tdp_api.h"
#include "../common/tmap.h"
#include "gemm_sm100.cuh"
#include "gemm_sm100_2cta.cuh"

namespace tdp {

namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld &&
           box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h = h * 1315423911u + k.inner;
    h = h * 1315423911u + k.outer;
    h = h * 1315423911u + k.ld;
    h = h * 1315423911u + k.box_inner * 1024 + k.box_outer;
    return h;
  }
};

// bf16 2-D row-major view [outer, inner] with leading dimension ld (elements), 128B swizzle.
}  // namespace

// (declared in common/tmap.h: shared with the attention kernels)
bool make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                  uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{base, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  auto fn = get_encode_fn();
  if (!fn) return false;
  // worker threads (autograd engine) may not have a current context yet: the driver entry point
  // below needs one.  Bind it once per thread (cudaSetDevice is legal during stream capture).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    int dev0 = 0;
    cudaGetDevice(&dev0);
    cudaSetDevice(dev0);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    // worker threads (autograd engine) may not have a current context yet: bind it and retry
    // (cudaSetDevice is legal during stream capture, cudaFree would not be)
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
    r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "[tdp] cuTensorMapEncodeTiled failed: CUresult=%d base=%p dims={%llu,%llu} ld=%llu "
            "box={%u,%u}\n",
            static_cast<int>(r), base, static_cast<unsigned long long>(inner),
            static_cast<unsigned long long>(outer), static_cast<unsigned long long>(ld), box_inner,
            box_outer);
    return false;
  }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return true;
}

namespace {

int g_num_sms = 0;

template <int BLOCK_N>
cudaError_t launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta_local,
                        const CUtensorMap& t_in, const CUtensorMap& t_aux, const GemmStoreMaps& sm,
                        const GemmParams& p, int grid, cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_sm100_kernel<BLOCK_N>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotalBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int threads = p.push_src != nullptr ? kGemmThreadsPush : kGemmThreads;
  gemm_bf16_sm100_kernel<BLOCK_N><<<grid, threads, S::kTotalBytes, stream>>>(
      ta, tb, ta_local, t_in, t_aux, sm, p);
  return cudaGetLastError();
}

template <int BLOCK_N, int EPI = 2>
cudaError_t launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& t_in,
                        const CUtensorMap& t_aux, const CUtensorMap& tc, const GemmParams& p,
                        int grid, cudaStream_t stream) {
  using S = Gemm2CtaSmem<BLOCK_N, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_sm100_2cta_kernel<BLOCK_N, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotalBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotalBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_bf16_sm100_2cta_kernel<BLOCK_N, EPI>, ta, tb, t_in, t_aux, tc, p);
}

}  // namespace

int gemm_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int launch_gemm_bf16(const GemmLaunch& g, cudaStream_t stream, const char** err) {
  static thread_local char msg[256];
  *err = msg;
  msg[0] = 0;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  if ((g.lda % 8) || (g.ldb % 8) || (reinterpret_cast<uintptr_t>(g.a) & 15) ||
      (reinterpret_cast<uintptr_t>(g.b) & 15)) {
    snprintf(msg, sizeof(msg), "gemm: operands must be 16-byte aligned with ld %% 8 == 0");
    return -1;
  }
  if ((g.ldc % 8) || (g.N % 8)) {
    snprintf(msg, sizeof(msg), "gemm: N and ldc must be multiples of 8 (N=%d ldc=%d)", g.N, g.ldc);
    return -1;
  }
  const int sms = gemm_num_sms();
  const int max_ctas = (g.max_ctas > 0 && g.max_ctas < sms) ? g.max_ctas : sms;
  const int mb = (g.M + kBlockM - 1) / kBlockM;

  int block_n = g.block_n;
  if (block_n != 128 && block_n != 256) {
    if (g.N <= 128) {
      block_n = 128;
    } else {
      // pick the tile width with the smaller (waves x tile cost); ties -> 256 (less smem traffic)
      const long t256 = static_cast<long>(mb) * ((g.N + 255) / 256);
      const long t128 = static_cast<long>(mb) * ((g.N + 127) / 128);
      const long c256 = ((t256 + max_ctas - 1) / max_ctas) * 256;
      const long c128 = ((t128 + max_ctas - 1) / max_ctas) * 128;
      // 128-wide tiles are L2->smem bandwidth bound (~1.05 PF measured vs ~1.45 PF for 256-wide):
      // only take them when they save more than ~15% of wave-quantised work
      block_n = (c128 * 115 < c256 * 100) ? 128 : 256;
    }
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.num_m_blocks = mb;
  p.num_n_blocks = (g.N + block_n - 1) / block_n;
  p.num_k_blocks = (g.K + kBlockK - 1) / kBlockK;
  p.a_mn_major = g.trans_a ? 1 : 0;   // A stored [K, M]: M contiguous
  p.b_mn_major = g.trans_b ? 0 : 1;   // B stored [K, N]: N contiguous
  p.C = g.c; p.ldc = g.ldc; p.c_fp32 = g.c_fp32; p.accumulate = g.accumulate;
  p.alpha = g.alpha;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(g.residual);
  p.ld_res = g.ld_res;
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(g.aux_in);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(g.aux_out);
  p.ld_aux = g.ld_aux;
  p.act = g.act;
  p.group_m = 8;
  p.comm_mode = g.comm_mode;
  p.rank = g.rank; p.world = g.world;
  p.rows_per_chunk = g.rows_per_chunk;
  p.chunk_flags = g.chunk_flags;
  p.flag_target = g.flag_target;
  for (int i = 0; i < kMaxPeers; ++i) p.peer_tile_counter[i] = g.peer_tile_counter[i];
  if (p.comm_mode != COMM_NONE) {
    if (g.world < 1 || g.world > kMaxPeers || g.rows_per_chunk % kBlockM != 0 ||
        g.rows_per_chunk * g.world != g.M) {
      snprintf(msg, sizeof(msg), "gemm(comm): M=%d must equal world(%d) x rows_per_chunk(%d), "
               "rows_per_chunk %% 128 == 0", g.M, g.world, g.rows_per_chunk);
      return -1;
    }
    // chunk-major tile order: a raster group never straddles two chunks
    const int bpc = g.rows_per_chunk / kBlockM;
    int gm = 8;
    while (bpc % gm) gm >>= 1;
    p.group_m = gm;
  }

  // grouped GEMM: operand extents cover all groups along whichever coordinate is shifted per group
  uint64_t a_m_ext = g.M, a_k_ext = g.K, b_n_ext = g.N, b_k_ext = g.K;
  if (g.grp_rows > 0) {
    if (g.grp_rows % kBlockM || g.M % g.grp_rows || g.K % kBlockK || g.comm_mode != 0 ||
        g.split_k > 1 || (g.grp_b_n != 0 && g.N % block_n)) {
      snprintf(msg, sizeof(msg), "gemm(grouped): rows/group %% 128, K %% 64 (and N %% tile for "
               "N-stacked B) must be 0; no collective / split-K modes");
      return -1;
    }
    const uint64_t G = g.M / g.grp_rows;
    p.grp_mblocks = g.grp_rows / kBlockM;
    p.grp_a_m = g.grp_a_m; p.grp_a_k = g.grp_a_k; p.grp_b_n = g.grp_b_n; p.grp_b_k = g.grp_b_k;
    p.grp_bias = g.grp_bias;
    if (g.grp_a_m != 0) a_m_ext = g.grp_rows;           // A's M extent is one group's rows
    if (g.grp_a_k != 0) a_k_ext = g.K * G;
    if (g.grp_b_n != 0) b_n_ext = g.N * G;
    if (g.grp_b_k != 0) b_k_ext = g.K * G;
  }
  CUtensorMap ta, tb;
  bool ok;
  if (!g.trans_a) ok = make_tmap_2d(&ta, g.a, a_k_ext, a_m_ext, g.lda, kBlockK, kBlockM);
  else            ok = make_tmap_2d(&ta, g.a, a_m_ext, a_k_ext, g.lda, 64, kBlockK);
  if (!ok) { snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(A) failed"); return -2; }
  if (g.trans_b)  ok = make_tmap_2d(&tb, g.b, b_k_ext, b_n_ext, g.ldb, kBlockK, block_n);
  else            ok = make_tmap_2d(&tb, g.b, b_n_ext, b_k_ext, g.ldb, 64, kBlockK);
  if (!ok) { snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(B) failed"); return -2; }

  // local (un-gathered) A shard for the all-gather variant: zero-copy, never waits
  CUtensorMap ta_local = ta;
  p.has_a_local = 0;
  if (p.comm_mode == COMM_AG_WAIT_A && g.a_local != nullptr && !g.trans_a) {
    if (!make_tmap_2d(&ta_local, g.a_local, g.K, g.rows_per_chunk, g.lda_local, kBlockK, kBlockM)) {
      snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(A local) failed");
      return -2;
    }
    p.has_a_local = 1;
  }
  if (p.comm_mode == COMM_AG_WAIT_A && g.push) {
    if (!p.has_a_local || g.lda_local != g.K || (static_cast<size_t>(g.rows_per_chunk) * g.K) % 8) {
      snprintf(msg, sizeof(msg), "gemm(ag push): needs a contiguous local shard");
      return -1;
    }
    static uint32_t* ticket = nullptr;      // one AG->GEMM at a time per process (it owns the GPU)
    if (ticket == nullptr) {
      if (cudaMalloc(&ticket, 64) != cudaSuccess) { snprintf(msg, sizeof(msg), "ticket alloc"); return -3; }
      cudaMemset(ticket, 0, 64);
    }
    p.push_src = reinterpret_cast<const uint4*>(g.a_local);
    p.push_vec = static_cast<size_t>(g.rows_per_chunk) * g.K * 2 / 16;
    p.push_mc = reinterpret_cast<char*>(g.push_mc);
    for (int r = 0; r < g.world; ++r) {
      p.push_dst[r] = reinterpret_cast<char*>(g.push_dst[r]);
      p.push_flag[r] = g.push_flag[r];
    }
    p.push_ticket = ticket;
  }

  // output maps: bf16 results leave through swizzled smem + TMA store (128 x 64 boxes)
  GemmStoreMaps sm;
  memset(&sm, 0, sizeof(sm));
  p.use_tma_store = 0;
  if (p.comm_mode == COMM_RS_SCATTER) {
    for (int r = 0; r < g.world; ++r) {
      const char* base = reinterpret_cast<const char*>(g.peer_out[r]) +
                         static_cast<size_t>(g.rank) * g.rows_per_chunk * g.ldc * 2;
      if (!make_tmap_2d(&sm.m[r], base, g.N, g.rows_per_chunk, g.ldc, kStoreCols, kBlockM)) {
        snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(peer C) failed");
        return -2;
      }
    }
    p.use_tma_store = 1;
  } else if (!g.c_fp32 && !g.accumulate && (reinterpret_cast<uintptr_t>(g.c) & 15) == 0) {
    if (!make_tmap_2d(&sm.m[0], g.c, g.N, g.M, g.ldc, kStoreCols, kBlockM)) {
      snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(C) failed");
      return -2;
    }
    p.use_tma_store = 1;
  }

  // row-wise epilogue operands travel through TMA as well (no strided 16-byte LSU traffic):
  //   input  = residual or the dGELU pre-activation, loaded per [128 x 64] sub-tile
  //   output = the pre-activation copy (aux_out), stored from the second staging buffer
  CUtensorMap t_in = ta, t_aux = ta;
  p.epi_in_tma = 0;
  p.epi_aux_tma = 0;
  if (p.use_tma_store && p.comm_mode != COMM_RS_SCATTER) {
    const bool has_res = g.residual != nullptr;
    const bool has_dgelu = (g.act == ACT_DGELU_TANH || g.act == ACT_DGELU_ERF) && g.aux_in != nullptr;
    if (has_res != has_dgelu) {       // exactly one row-wise input
      const void* base = has_res ? g.residual : g.aux_in;
      const int ld = has_res ? g.ld_res : g.ld_aux;
      if ((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0 &&
          make_tmap_2d(&t_in, base, g.N, g.M, ld, kStoreCols, kBlockM))
        p.epi_in_tma = has_res ? 1 : 2;
    }
    if (g.aux_out != nullptr && p.epi_in_tma == 0 &&
        (reinterpret_cast<uintptr_t>(g.aux_out) & 15) == 0 && g.ld_aux % 8 == 0 &&
        make_tmap_2d(&t_aux, g.aux_out, g.N, g.M, g.ld_aux, kStoreCols, kBlockM))
      p.epi_aux_tma = 1;
  }

  // opt-in: two decoupled epilogue groups (see gemm_sm100.cuh); plain TMA-store GEMMs only
  static const bool env_epi_split = [] { const char* v = getenv("TDP_GEMM_EPI"); return v && !strcmp(v, "split"); }();
  p.epi_split = (env_epi_split && p.use_tma_store && p.comm_mode == COMM_NONE && g.split_k <= 1) ? 1 : 0;

  // split-K (weight-gradient shapes: few output tiles, very long K): partials are added with
  // vector atomics into an fp32 output that the caller zero-initialised (or accumulates into)
  p.split_k = 1;
  p.k_blocks_per_split = p.num_k_blocks;
  if (g.split_k > 1) {
    if (!g.c_fp32 || g.comm_mode != 0 || g.bias || g.residual || g.aux_in || g.aux_out ||
        g.act != 0) {
      snprintf(msg, sizeof(msg), "gemm: split_k needs a plain fp32 output (no fused epilogue)");
      return -1;
    }
    p.use_tma_store = 0;
  }

  const long tiles = static_cast<long>(p.num_m_blocks) * p.num_n_blocks;
  int grid = static_cast<int>(tiles < max_ctas ? tiles : max_ctas);
  if (g.split_k > 1) {
    // stream-K: equal contiguous shares of the flattened (tile, k-block) space, one per CTA
    const long total = tiles * p.num_k_blocks;
    long ctas = total / 8 < max_ctas ? total / 8 : max_ctas;      // >= 8 k-blocks per CTA
    if (ctas < 1) ctas = 1;
    const long share = (total + ctas - 1) / ctas;
    p.split_k = 2;                                    // flag: stream-K on
    p.k_blocks_per_split = static_cast<int>(share);
    grid = static_cast<int>((total + share - 1) / share);
  }
  // 2-CTA (cta_group::2, 256 x 256 per CTA pair) path for plain bf16-output GEMMs
  // auto (cta_group == 0): CTA pairs pay off when the main loop dominates (long K: B is fetched
  // once per pair, 6 smem stages) -- measured on B200: +8..12 % at K >= 2304, neutral at K = 768,
  // slower with the heavy GELU epilogues.  TDP_GEMM_2CTA=0/1 forces the choice.
  static const int env_2cta = [] { const char* v = getenv("TDP_GEMM_2CTA"); return v ? (v[0] == '1' ? 1 : 0) : -1; }();
  bool auto_2cta = g.K >= 2048 && g.act == ACT_NONE && g.aux_out == nullptr;
  // short-K products with a heavy fused epilogue (GELU + pre-activation output, GELU' x gradient,
  // residual): the pair kernel's variant with four operand stages and a six-deep epilogue staging
  // ring (gemm_sm100_2cta.cuh).  TDP_GEMM_EPIRING=0 keeps them on the 1-CTA kernel.
  static const int env_ring = [] { const char* v = getenv("TDP_GEMM_EPIRING"); return v ? (v[0] == '1' ? 1 : 0) : -1; }();
  const bool heavy_epi = (g.act != ACT_NONE || g.aux_out != nullptr || p.epi_in_tma != 0) &&
                         g.K <= 1536 && g.N >= 512 && g.M >= 1024 && g.block_n != 128;
  const bool epi_ring = env_ring >= 0 ? (env_ring == 1 && heavy_epi) : heavy_epi;
  if (epi_ring && g.cta_group != 1) auto_2cta = true;
  if (env_2cta >= 0 && !epi_ring) auto_2cta = env_2cta == 1;
  if (g.grp_rows == 0 && (g.cta_group == 2 || (g.cta_group == 0 && auto_2cta)) && p.comm_mode == COMM_NONE &&
      p.split_k == 1 && p.use_tma_store && g.N > 128 && g.M > 128) {
    GemmParams q = p;
    q.num_m_blocks = (g.M + 2 * kBlockM - 1) / (2 * kBlockM);
    const int clusters = (max_ctas & ~1) / 2;
    // pair tile 256 x 256, or 256 x 128 when the wide tiles cannot occupy every CTA pair
    int bn2 = g.block_n == 128 ? 128 : 256;
    if (g.block_n == 0 && !epi_ring &&
        static_cast<long>(q.num_m_blocks) * ((g.N + 255) / 256) < clusters) bn2 = 128;
    q.num_n_blocks = (g.N + bn2 - 1) / bn2;
    q.group_m = 4;
    CUtensorMap tb2 = tb;
    if (g.trans_b && !make_tmap_2d(&tb2, g.b, g.K, g.N, g.ldb, kBlockK, bn2 / 2)) {
      snprintf(msg, sizeof(msg), "gemm: cuTensorMapEncodeTiled(B, 2cta) failed");
      return -2;
    }
    const long tiles2 = static_cast<long>(q.num_m_blocks) * q.num_n_blocks;
    const int ctas = static_cast<int>(tiles2 < clusters ? tiles2 : clusters) * 2;
    cudaError_t e2 = bn2 == 256
        ? (epi_ring ? launch_2cta<256, 6>(ta, tb2, t_in, t_aux, sm.m[0], q, ctas, stream)
                    : launch_2cta<256>(ta, tb2, t_in, t_aux, sm.m[0], q, ctas, stream))
        : launch_2cta<128>(ta, tb2, t_in, t_aux, sm.m[0], q, ctas, stream);
    if (e2 != cudaSuccess) {
      snprintf(msg, sizeof(msg), "gemm 2cta launch: %s", cudaGetErrorString(e2));
      return static_cast<int>(e2);
    }
    return 0;
  }

  cudaError_t e = (block_n == 256)
                      ? launch_impl<256>(ta, tb, ta_local, t_in, t_aux, sm, p, grid, stream)
                      : launch_impl<128>(ta, tb, ta_local, t_in, t_aux, sm, p, grid, stream);
  if (e != cudaSuccess) {
    snprintf(msg, sizeof(msg), "gemm launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace tdp
