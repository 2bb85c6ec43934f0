// PTX helpers shared by the tcgen05 kernels (sm_100a): mbarrier, cp.async(.bulk), tcgen05.mma/commit/ld, UMMA
// shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout), tf32 hi/lo split.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace p3d {
namespace tc {

constexpr int kM = 128;
constexpr int kProducers = 128;
constexpr int kThreads = 192;  // 4 producer/epilogue warps + 1 MMA warp + 1 weight-TMA warp

__host__ __device__ constexpr int kc_of(int cin) { return cin < 16 ? cin : 16; }  // channels per pipeline stage

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef P3D_MBAR_HINT_NS
#define P3D_MBAR_HINT_NS 20000  // suspend-time hint of mbarrier.try_wait (ns); build with -DP3D_MBAR_HINT_NS=.. to experiment
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or ~20 us pass)
  // instead of polling - with the short default limit the waiting warps executed ~70 % of all instructions of the
  // sparse-conv kernels (ncu: 1400 warp instructions per pipeline use, issue slots 50 % busy).
  uint32_t done;
  int spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(static_cast<uint32_t>(P3D_MBAR_HINT_NS))
        : "memory");
    if (!done && ++spins > 200000) __trap();  // seconds: a protocol bug must not hang the GPU
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; LBO = distance between the two 16-byte
// K-chunks of one MMA (K = 8 tf32), SBO = distance between 8-row groups (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;  // descriptor version 1 (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float r = x - hi;  // exact in fp32
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}


// split-K factor of the tensor-core sparse conv (taps spread over this many CTAs per 128-row tile)
inline int splits_for(int Cout) {
  // tuning hooks (not a public knob): P3D_SPLITS_128 / P3D_SPLITS_64 / P3D_SPLITS_32 override the defaults
  static const int s128 = getenv("P3D_SPLITS_128") ? atoi(getenv("P3D_SPLITS_128")) : 2;
  static const int s64 = getenv("P3D_SPLITS_64") ? atoi(getenv("P3D_SPLITS_64")) : 2;
  static const int s32 = getenv("P3D_SPLITS_32") ? atoi(getenv("P3D_SPLITS_32")) : 1;
  const int v = Cout >= 128 ? s128 : (Cout >= 64 ? s64 : s32);
  return v < 1 ? 1 : (v > 8 ? 8 : v);
}

}  // namespace tc
}  // namespace p3d
