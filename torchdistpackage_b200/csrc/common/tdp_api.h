// C++ launcher API between the torch bindings (bindings.cpp, compiled by g++) and the
// CUDA translation units (compiled by nvcc, no torch headers -> fast rebuilds).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace tdp {

constexpr int kApiMaxPeers = 8;

// ------------------------------------------------------------------ GEMM (gemm/gemm.cu)
struct GemmLaunch {
  // op(A)[M,K] @ op(B)[K,N]; bf16 operands
  const void* a;  // trans_a==0: [M, lda] row major ; trans_a==1: [K, lda] row major (lda >= M)
  const void* b;  // trans_b==0: [K, ldb] row major ; trans_b==1: [N, ldb] row major (ldb >= K)
  int lda, ldb;
  int trans_a, trans_b;
  int M, N, K;
  void* c;        // [M, ldc] bf16 or fp32
  int ldc;
  int c_fp32;
  int accumulate;
  float alpha;
  const void* bias;      // bf16 [N]
  const void* residual;  // bf16 [M, ld_res]
  int ld_res;
  const void* aux_in;    // bf16 [M, ld_aux]
  void* aux_out;         // bf16 [M, ld_aux]
  int ld_aux;
  int act;               // GemmAct
  int split_k;           // >1: split K over CTAs (fp32 output, atomically accumulated)
  int block_n;           // 0 = auto, 128 or 256
  int cta_group;         // 0 = auto, 1 = one CTA per tile, 2 = CTA pairs (cta_group::2, 256x256)
  int max_ctas;          // 0 = all SMs (fused collectives reserve SMs for the comm CTAs)
  // fused collective hooks (see gemm_sm100.cuh)
  int comm_mode;
  int rank, world;
  int rows_per_chunk;
  const void* a_local;   // AG mode: this rank's un-gathered shard [rows_per_chunk, lda_local]
  int lda_local;
  uint32_t* chunk_flags;
  uint32_t flag_target;
  void* peer_out[kApiMaxPeers];
  uint32_t* peer_tile_counter[kApiMaxPeers];
  // grouped GEMM (MoE experts stacked along M): rows_per_group % 128 == 0 rows of C form a group;
  // group g reads its operands at coordinate offsets g * {a_m, a_k, b_n, b_k} (elements) and adds
  // bias + g * bias_stride.  M, N, K describe ONE group's product except M = all groups' rows.
  int grp_rows;          // 0 = plain GEMM
  int grp_a_m, grp_a_k, grp_b_n, grp_b_k, grp_bias;
  // AG mode, push folded into the GEMM: a_local (contiguous) is copied by an extra warp of every
  // CTA to push_dst[r] (+ my own slot) / push_mc, then push_flag[r] = flag_target is published
  int push;
  void* push_dst[kApiMaxPeers];
  void* push_mc;
  uint32_t* push_flag[kApiMaxPeers];
};
// returns 0 on success, otherwise a cudaError_t / negative code; `err` receives a message
int launch_gemm_bf16(const GemmLaunch& g, cudaStream_t stream, const char** err);
int gemm_num_sms();

// ------------------------------------------------------------------ collectives (coll/*.cu)
struct SymmPeers {
  int rank, world;
  void* buf[kApiMaxPeers];        // peer-mapped base pointers of the symmetric buffer
  uint32_t* signal[kApiMaxPeers]; // peer-mapped signal pads (uint32 words)
  void* mc_buf;                   // multicast mapping of the buffer (nullptr if unavailable)
};

// device-side barrier across the ranks of a symmetric group (signal pad slot `slot`)
void launch_symm_barrier(const SymmPeers& s, int slot, cudaStream_t stream);

// in-place all-reduce of `numel` elements at byte offset `offset` of the symmetric buffer.
// dtype: 0 = bf16, 1 = fp32.  scale is applied to the sum (1/N for AVG).
// algo: 0 = auto, 1 = one-shot p2p, 2 = two-shot p2p, 3 = NVLS multimem two-shot, 4 = one-shot NVLS.
// one_shot_scratch: kOneShotScratchBytes of device memory private to this symmetric buffer (the
// one-shot latency path for messages up to that size; nullptr disables it).
constexpr size_t kOneShotScratchBytes = 256 * 1024;
void launch_all_reduce(const SymmPeers& s, size_t offset, size_t numel, int dtype, float scale,
                       int algo, int max_ctas, cudaStream_t stream, void* one_shot_scratch = nullptr);

// reduce-scatter: rank r ends up with the reduced slice r of [world * slice_numel] at `offset`
// written to `out` (bf16/fp32 selectable; out_fp32 accumulates into fp32 master grads).
// reduce-scatter -> AdamW on this rank's 1/N shard -> all-gather of the new bf16 parameters in ONE
// kernel (collectives.cu).  `grad` / `param`: symmetric buffers of the same group; offsets in
// bytes; n_vec = 16-byte vectors (8 bf16) in the bucket payload; master / exp_avg / exp_avg_sq: this
// rank's fp32 shard, ceil(n_vec / world) * 8 floats each; hyper: device floats {step, lr}.
struct FusedAdamLaunch {
  size_t grad_offset, param_offset, n_vec;
  float* master;
  float* exp_avg;
  float* exp_avg_sq;
  float beta1, beta2, eps, weight_decay;
  int adamw_mode;
  const float* hyper;
  float grad_scale;
};
void launch_fused_rs_adamw_ag(const SymmPeers& grad, const SymmPeers& param,
                              const FusedAdamLaunch& a, int use_mc, int write_back_grad,
                              int max_ctas, cudaStream_t stream);
void launch_reduce_scatter(const SymmPeers& s, size_t offset, size_t slice_numel, int dtype,
                           float scale, void* out, int out_fp32, int accumulate_out, int use_mc,
                           int max_ctas, cudaStream_t stream);

// all-gather: every rank contributes `slice_numel` elements from `src` (local memory, same dtype)
// into slot `rank` of the symmetric buffer region [world * slice_numel] on every rank.
void launch_all_gather(const SymmPeers& s, size_t offset, size_t slice_numel, int elem_bytes,
                       const void* src, int use_mc, int max_ctas, cudaStream_t stream);

// all-gather with per-chunk completion flags, used by the fused all-gather->GEMM:
// pushes this rank's slice to all peers then bumps chunk_flag[rank] on every peer to `flag_value`.
void launch_all_gather_signal(const SymmPeers& s, size_t offset, size_t slice_bytes,
                              const void* src, size_t flag_word_offset, uint32_t flag_value,
                              int use_mc, int max_ctas, cudaStream_t stream);

// reduce the `world` partial slices that peers scattered into my staging buffer
// (layout [world][rows][ld] bf16 at `offset`), waiting until counters[src] >= target for all src.
// out = sum (+bias) (+residual); optionally broadcast the result to all ranks (all-reduce mode)
struct RsReduceLaunch {
  size_t offset;            // byte offset of the staging region in the symmetric buffer
  int rows, cols, ld;       // slice geometry (bf16 elements)
  size_t counter_word_offset;  // signal pad word index of counters[world]
  uint32_t counter_target;
  const void* bias;         // bf16 [cols] or null
  const void* residual;     // bf16 [rows, ld_res] or null
  int ld_res;
  void* out;                // bf16 [rows, ld_out] local output (RS)
  int ld_out;
  int broadcast;            // 1: write rows into slot `rank` of out region on every rank (AR)
  size_t bcast_offset;      // byte offset (symmetric buffer) of the [world*rows, ld_out] region
};
void launch_rs_reduce(const SymmPeers& s, const RsReduceLaunch& r, int use_mc, int max_ctas,
                      cudaStream_t stream);

// MoE all-to-all: scatter rows of `src` (bf16 [n_rows, hidden]) to peer buffers.
//   dst_rank[i], dst_row[i] give the destination of row i (dst_row < 0: dropped).
void launch_a2a_scatter_rows(const SymmPeers& s, size_t offset, const void* src, int n_rows,
                             int hidden, const int* dst_rank, const int* dst_row,
                             cudaStream_t stream);
// gather rows back: out[i] = scale[i] * peer(src_rank[i]).buf[offset + src_row[i]]
void launch_a2a_gather_rows(const SymmPeers& s, size_t offset, void* out, int n_rows, int hidden,
                            const int* src_rank, const int* src_row, const float* scale,
                            int accumulate, cudaStream_t stream);

// ------------------------------------------------------------------ fused ops (fused/*.cu)
struct AdamWLaunch {
  void* param;        // bf16 or fp32 model params (flat)
  float* master;      // fp32 master copy or nullptr (then param must be fp32)
  const void* grad;   // bf16 or fp32 flat grads
  float* exp_avg;
  float* exp_avg_sq;
  size_t numel;
  int param_bf16, grad_bf16;
  float lr, beta1, beta2, eps, weight_decay;
  float bias_correction1, bias_correction2;
  float grad_scale;        // multiplied into grad (unscale / clip coefficient)
  const float* grad_scale_ptr;  // optional device scalar multiplied in as well
  void* param_copy_out;    // optional second bf16 destination (e.g. symmetric all-gather slot)
  int adamw_mode;          // 1: decoupled weight decay (AdamW), 0: L2 added to the gradient (Adam)
  const float* hyper;      // optional device floats {step, lr}: overrides lr / bias corrections
};
void launch_adamw(const AdamWLaunch& a, cudaStream_t stream);

void launch_ema_update(float* ema, const void* param, int param_bf16, size_t numel, float decay,
                       cudaStream_t stream);
// multi-tensor variants: arrays of device pointers / sizes living in device memory
void launch_ema_update_multi(void* const* ema_ptrs, const void* const* param_ptrs,
                             const int64_t* numels, const int* ema_dtypes, const int* param_dtypes,
                             int n_tensors, float decay, cudaStream_t stream);

// sum of squares of a flat buffer into *out (fp32, atomically accumulated; caller zeroes)
void launch_sumsq_multi(const void* const* ptrs, const int64_t* numels, const int* dtypes,
                        int n_tensors, float* out, cudaStream_t stream);
void launch_scale_multi(void* const* ptrs, const int64_t* numels, const int* dtypes, int n_tensors,
                        float scale, const float* scale_ptr, cudaStream_t stream);
void launch_sumsq(const void* x, int dtype, size_t numel, float* out, cudaStream_t stream);
void launch_scale_(void* x, int dtype, size_t numel, float scale, const float* scale_ptr,
                   cudaStream_t stream);
void launch_cast_copy(void* dst, int dst_dtype, const void* src, int src_dtype, size_t numel,
                      float scale, cudaStream_t stream);

// strided copy of rows (row_bytes % 16 == 0, contiguous rows) indexed by (i0,i1,i2); byte strides
// dst[idx[r], :] += src[r, :] for bf16 rows of row_bytes (multiple of 16) bytes (fused/layout.cu)
void launch_rows_scatter_add_bf16(const void* src, void* dst, const long* idx, int n_rows,
                                  int row_bytes, long dst_rows, cudaStream_t stream);
void launch_rows_copy3(const void* const* src, void* const* dst, int n0, int n1, int n2,
                       int row_bytes, long s0, long s1, long s2, long d0, long d1, long d2,
                       cudaStream_t stream);
void launch_rows_copy(const void* src, void* dst, int n0, int n1, int n2, int row_bytes, long s0,
                      long s1, long s2, long d0, long d1, long d2, cudaStream_t stream);

// LayerNorm over the last dim of bf16 [rows, cols]; saves mean / rstd (fp32) for backward
void launch_layernorm_fwd(const void* x, const void* residual, const void* gamma, const void* beta,
                          void* y, void* resid_out, float* mean, float* rstd, int rows, int cols,
                          float eps, cudaStream_t stream);
void launch_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                          const float* rstd, void* dx, const void* dresid, float* dgamma_partial,
                          float* dbeta_partial, int rows, int cols, int n_partial,
                          cudaStream_t stream);
void launch_colsum_partial_reduce(const float* partial, int n_partial, int cols, void* out,
                                  int out_bf16, cudaStream_t stream);
void launch_colsum_partial_reduce2(const float* partial, int n_partial, int cols, void* out0,
                                   void* out1, int out_bf16, cudaStream_t stream);
// column sum of bf16 [rows, cols] (bias gradient)
void launch_colsum(const void* x, int rows, int cols, int ld, float* scratch, void* out,
                   int out_bf16, cudaStream_t stream);

// fused softmax cross entropy over bf16 logits [rows, vocab] (ld >= vocab): writes per-row loss
// (fp32) and overwrites logits with dlogits * grad_scale (bf16) in one pass.
void launch_cross_entropy_fwd_bwd(void* logits, int rows, int vocab, int ld, const int64_t* target,
                                  float* loss, float grad_scale, int ignore_index,
                                  cudaStream_t stream);

// ------------------------------------------------------------------ attention (attn/*.cu)
// Flash-attention forward, bf16, head_dim 64.  q / k / v / o are token matrices: row b*T + t,
// row stride ld_* elements, head h at columns [h*D, (h+1)*D) from the given base pointer (so q, k
// and v may be three column windows of one packed qkv projection).  lse: fp32 [B, H, T] or null.
struct AttnFwdLaunch {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  int B, T, H, D;
  int ld_q, ld_k, ld_v, ld_o;
  int causal;
  float scale;
};
int launch_attn_fwd(const AttnFwdLaunch& a, cudaStream_t stream, const char** err);

// Flash-attention backward (same token-matrix contract), two kernels launched back to back:
//   1. attn_bwd_dq_sm100.cu : delta = rowsum(dO * O) -> `delta` (fp32 [B, H, T] scratch) and dQ
//   2. attn_bwd_sm100.cu    : dK, dV (reads lse and delta)
// lse: forward's [B, H, T]; o: forward output; dq / dk / dv: bf16 token matrices (typically the
// three column windows of one packed dqkv gradient).  No atomics, no fp32 accumulator.
struct AttnBwdLaunch {
  const void* q;
  const void* k;
  const void* v;
  const void* o;
  const void* d_o;
  const float* lse;
  float* delta;
  void* dq;
  void* dk;
  void* dv;
  int B, T, H, D;
  int ld_q, ld_k, ld_v, ld_o, ld_do, ld_dq, ld_dk, ld_dv;
  int causal;
  float scale;
};
int launch_attn_bwd_dq(const AttnBwdLaunch& a, cudaStream_t stream, const char** err);
int launch_attn_bwd_dkv(const AttnBwdLaunch& a, cudaStream_t stream, const char** err);
// both, in order
int launch_attn_bwd(const AttnBwdLaunch& a, cudaStream_t stream, const char** err);

// ------------------------------------------------------------------ native symmetric memory
// (symm/symm_vmm.cpp): CUDA VMM allocations exported as POSIX fds, peer mapping, NVLS multicast
bool vmm_granularity(int device, int num_devices, uint64_t* gran, std::string& err);
bool vmm_alloc(uint64_t size, int device, uint64_t* handle, uint64_t* ptr, int* fd, std::string& err);
bool vmm_import(int fd, uint64_t size, int device, uint64_t* handle, uint64_t* ptr, std::string& err);
bool mc_create(uint64_t size, int num_devices, uint64_t* handle, int* fd, std::string& err);
bool mc_import(int fd, uint64_t* handle, std::string& err);
bool mc_add_device(uint64_t mc_handle, int device, std::string& err);
bool mc_bind_and_map(uint64_t mc_handle, uint64_t mem_handle, uint64_t size, int device,
                     uint64_t* mc_ptr, std::string& err);
bool vmm_unmap_release(uint64_t handle, uint64_t ptr, uint64_t size, std::string& err);

}  // namespace tdp
