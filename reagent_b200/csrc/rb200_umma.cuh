// reagent_b200 -- tcgen05 / TMEM / mbarrier / bulk-copy PTX helpers shared by the Blackwell
// tensor-core kernels (rb200_tc_gemm.cu, rb200_dqn_tc.cu).  sm_100a only.
#pragma once
#include "rb200_common.cuh"

namespace rb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte
// offset >> 4, [46,48) version = 1 (sm_100), [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, K-major A and B.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 2u << 7;                      // a_format = TF32
  d |= 2u << 10;                     // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;     // n_dim
  d |= (uint32_t)(M >> 4) << 24;     // m_dim
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (long long it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (it > 400000LL) __trap();  // ~2 s (a failed try_wait blocks a few us): never spin forever on a protocol bug
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
      smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  uint32_t h;
  h = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u; hi.x = __uint_as_float(h); lo.x = v.x - hi.x;
  h = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u; hi.y = __uint_as_float(h); lo.y = v.y - hi.y;
  h = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u; hi.z = __uint_as_float(h); lo.z = v.z - hi.z;
  h = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u; hi.w = __uint_as_float(h); lo.w = v.w - hi.w;
}

// ---- mbarrier transaction + 1-D bulk copy (TMA engine, no tensor map) ----
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// global -> shared, `bytes` % 16 == 0, both addresses 16 B aligned; completion is signalled on
// `bar` as a transaction count.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// one lane of a converged warp (the same lane on every call)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

}  // namespace rb200
