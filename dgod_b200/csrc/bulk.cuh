// PTX wrappers shared by the bulk-copy (TMA engine) kernels: mbarriers, cp.async.bulk loads / stores /
// reductions, L2 eviction-priority policies.  sm_100a only.
#pragma once
#include "common.cuh"

namespace dgod {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on try_wait (no suspend-time hint: with a hint ptxas emits a NANOSLEEP back-off loop whose wake-up
// granularity adds ~1 us to every wait that actually blocks — measured on the owner-computes backward's ring).
#ifndef DGOD_MBAR_WAIT_HINT
#define DGOD_MBAR_WAIT_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
#if DGOD_MBAR_WAIT_HINT
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
// The same on 32-bit shared-window addresses computed once (smem_opaque): with generic pointers to static shared
// variables the compiler rematerialises the window base (S2R SR_CgaCtaId + LEA) in front of every barrier operation.
__device__ __forceinline__ unsigned smem_opaque(const void* p) {
  unsigned a;
  asm volatile("mov.u32 %0, %1;\n" : "=r"(a) : "r"(smem_u32(p)));
  return a;
}
__device__ __forceinline__ void mbar_wait_s(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4s(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32s(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_bf16s(unsigned a) {       // bf16 -> fp32 (exact)
  unsigned short u;
  asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(u) : "r"(a));
  return __uint_as_float((unsigned)u << 16);
}
__device__ __forceinline__ unsigned lds_u8s(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s16s(unsigned a) {
  short v;
  asm volatile("ld.shared.s16 %0, [%1];\n" : "=h"(v) : "r"(a));
  return (int)v;
}
__device__ __forceinline__ unsigned long long lds_u64s(unsigned a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u4s(unsigned a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f4s(unsigned a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Pure polling (test_wait never suspends the thread): for waits on the critical path of a producer/consumer ring.
__device__ __forceinline__ void mbar_spin(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SPIN_%=:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SDONE_%=;\n"
      "bra SPIN_%=;\n"
      "SDONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
template <typename T> __device__ __forceinline__ void bulk_reduce_add(void* gmem_dst, const void* smem_src, unsigned bytes);
template <> __device__ __forceinline__ void bulk_reduce_add<float>(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
template <> __device__ __forceinline__ void bulk_reduce_add<__nv_bfloat16>(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// L2 eviction-priority hints: streamed-once data (plans, gradient blocks, pooled output) is marked
// evict_first so that it does not push the re-used maps (features / gradient maps of the image in
// flight, marked evict_last) out of the L2.
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                               unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* gmem_dst, const void* smem_src, unsigned bytes, unsigned long long policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
template <typename T> __device__ __forceinline__ void bulk_reduce_add_hint(void* gmem_dst, const void* smem_src, unsigned bytes,
                                                                           unsigned long long policy);
template <> __device__ __forceinline__ void bulk_reduce_add_hint<float>(void* gmem_dst, const void* smem_src, unsigned bytes,
                                                                        unsigned long long policy) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f32 [%0], [%1], %2, %3;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
template <> __device__ __forceinline__ void bulk_reduce_add_hint<__nv_bfloat16>(void* gmem_dst, const void* smem_src, unsigned bytes,
                                                                                unsigned long long policy) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.noftz.bf16 [%0], [%1], %2, %3;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void st_zero16_hint(void* p, unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %1, %1, %1}, %2;\n" ::"l"(p), "r"(0u), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(N) : "memory"); }


}  // namespace dgod
