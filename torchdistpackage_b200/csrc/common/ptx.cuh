// sm_100a inline-PTX wrappers used by every kernel in this tree.
//   tcgen05.{alloc,dealloc,mma,commit,ld,fence}, cp.async.bulk.tensor (TMA),
//   mbarrier, multimem.{ld_reduce,st,red}, sys-scope acquire/release, misc.
// Hand written for B200 (compute_100a); nothing here is portable to older parts.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tdp {

#define TDP_DEVICE __device__ __forceinline__

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
TDP_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

TDP_DEVICE uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

TDP_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

TDP_DEVICE uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Watchdog for cross-GPU spin waits: a peer that never arrives must become an error
// (cudaErrorLaunchFailure on this rank) instead of a hung box.  ~10 s.
constexpr uint64_t kSpinTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;
struct SpinWatchdog {
  uint64_t t0;
  uint32_t polls;
  TDP_DEVICE SpinWatchdog() : t0(0), polls(0) {}
  // call once per failed poll; `what` / a / b are printed when the wait times out
  TDP_DEVICE void tick(const char* what, int a, int b) {
    if ((++polls & 0x3FFu) != 0u) return;
    const uint64_t now = globaltimer_ns();
    if (t0 == 0) { t0 = now; return; }
    if (now - t0 > kSpinTimeoutNs) {
      printf("[tdp] spin wait timed out: %s (%d, %d) block %d\n", what, a, b, blockIdx.x);
      __trap();
    }
  }
};

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
TDP_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

TDP_DEVICE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

TDP_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

TDP_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

TDP_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// non-blocking phase test (try_wait may suspend the thread for a system-dependent time when the
// phase is not complete -- wrong tool for a loop that polls several barriers)
TDP_DEVICE bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

TDP_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  SpinWatchdog wd;
  while (!mbar_try_wait(bar, parity)) {
    wd.tick("mbarrier", static_cast<int>(smem_u32(bar)), static_cast<int>(parity));
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
TDP_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes)
TDP_DEVICE void tma_load_2d(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0,
                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

TDP_DEVICE void tma_load_2d_hint(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0,
                                 int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(cache_hint)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group)
TDP_DEVICE void tma_store_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}
TDP_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
TDP_DEVICE void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
TDP_DEVICE void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// generic-proxy writes -> visible to async proxy (TMA / tcgen05 reads of smem or global)
TDP_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
TDP_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
TDP_DEVICE void tmem_alloc(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_holder)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
TDP_DEVICE void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
TDP_DEVICE void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
TDP_DEVICE void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/fp16 inputs, fp32 accumulate, one CTA.
TDP_DEVICE void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued tcgen05.mma of this thread arrive (once) on the mbarrier when complete
TDP_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i)
TDP_DEVICE void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
TDP_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   start address, LBO, SBO are encoded >> 4.  layout_type 2 == SWIZZLE_128B.
TDP_DEVICE uint64_t make_umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t umma_m, uint32_t umma_n,
                                                           uint32_t a_mn_major,
                                                           uint32_t b_mn_major) {
  return (1u << 4)                   // D format  = F32
         | (1u << 7)                 // A format  = BF16
         | (1u << 10)                // B format  = BF16
         | ((a_mn_major & 1u) << 15) // A major
         | ((b_mn_major & 1u) << 16) // B major
         | ((umma_n >> 3) << 17)     // N / 8
         | ((umma_m >> 4) << 24);    // M / 16
}

// ----------------------------------------------------------------------------------------------
// system-scope synchronisation (peer GPUs over NVLink)
// ----------------------------------------------------------------------------------------------
TDP_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TDP_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
TDP_DEVICE uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
TDP_DEVICE void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TDP_DEVICE void red_add_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TDP_DEVICE void red_add_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TDP_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
TDP_DEVICE void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TDP_DEVICE void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// 16-byte peer/global accesses that bypass L1 (streaming)
TDP_DEVICE uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
TDP_DEVICE uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
TDP_DEVICE void st_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// NVLS multicast (multimem.*): the address must be inside a multicast mapping.
// ----------------------------------------------------------------------------------------------
// in-switch reduction of 8 bf16 values (as 4 x bf16x2) held at the same offset on every member
TDP_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc_ptr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_ptr)
               : "memory");
  return v;
}
TDP_DEVICE float4 multimem_ld_reduce_f32x4(const void* mc_ptr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_ptr)
               : "memory");
  return v;
}
// broadcast store of 16 bytes to every member of the multicast group
TDP_DEVICE void multimem_st_v4(void* mc_ptr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// broadcast reduction (every member's copy += v), 8 bf16
TDP_DEVICE void multimem_red_bf16x8(void* mc_ptr, const uint4& v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
TDP_DEVICE void multimem_red_add_u32(void* mc_ptr, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_ptr), "r"(v)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// numeric helpers
// ----------------------------------------------------------------------------------------------
TDP_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
TDP_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// hardware tanh (one MUFU op, ~2^-11 relative error: far below bf16 resolution)
TDP_DEVICE float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
TDP_DEVICE float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float u = x * fmaf(k01, x * x, k0);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
TDP_DEVICE float dgelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(k01, x2, k0));
  const float du = fmaf(3.f * k01, x2, k0);
  const float w = (0.5f * x) * fmaf(-t, t, 1.f);          // 0.5 x sech^2(u)
  return fmaf(w, du, fmaf(0.5f, t, 0.5f));
}
// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 do two fp32 lanes per issue slot).
// The GEMM epilogues are issue-bound, so the element-wise math runs on register pairs.
using f32x2 = unsigned long long;
TDP_DEVICE f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
TDP_DEVICE void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
TDP_DEVICE f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
TDP_DEVICE f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
TDP_DEVICE f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
TDP_DEVICE f32x2 splat2(float x) { return pk2(x, x); }
// bf16x2 word -> two fp32 lanes (a shift and a mask)
TDP_DEVICE f32x2 bf16x2_to_f32x2(uint32_t u) {
  return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
TDP_DEVICE uint32_t f32x2_to_bf16x2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
TDP_DEVICE f32x2 tanh_approx2(f32x2 u) {
  float a, b;
  upk2(u, a, b);
  return pk2(tanh_approx(a), tanh_approx(b));
}
TDP_DEVICE f32x2 gelu_tanh2(f32x2 x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const f32x2 u = mul2(x, fma2(splat2(k01), mul2(x, x), splat2(k0)));
  const f32x2 hx = mul2(x, splat2(0.5f));
  return fma2(hx, tanh_approx2(u), hx);
}
// returns g * gelu'(x)
TDP_DEVICE f32x2 dgelu_tanh2(f32x2 x, f32x2 g) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const f32x2 x2 = mul2(x, x);
  const f32x2 t = tanh_approx2(mul2(x, fma2(splat2(k01), x2, splat2(k0))));
  const f32x2 du = fma2(splat2(3.f * k01), x2, splat2(k0));
  const f32x2 sech = fma2(mul2(t, splat2(-1.f)), t, splat2(1.f));
  const f32x2 w = mul2(mul2(x, splat2(0.5f)), sech);                 // 0.5 x sech^2(u)
  return mul2(g, fma2(w, du, fma2(splat2(0.5f), t, splat2(0.5f))));
}
TDP_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }
TDP_DEVICE float dgelu_erf(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.7071067811865476f));
  float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

}  // namespace tdp
