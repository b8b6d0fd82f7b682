// gw_tc_ptx.cuh -- PTX wrappers shared by the tensor-core chain kernels (mbarrier, bulk copy, tcgen05, UMMA descriptors).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gw {

// ------------------------------------------------------------------------------------------------------------------
// PTX wrappers (syntax checked against cute/arch/{mma_sm100_umma,copy_sm100,tmem_allocator_sm100}.hpp)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  After ~4 s the thread records which barrier it was waiting on in the
// (host-mapped) status block -- word 0 bit 1, words 1..5 = barrier byte offset, parity, thread, block, role tag -- and traps.
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int32_t* status) {
  if (status) {
    extern __shared__ __align__(1024) uint8_t smem_dbg[];
    atomicOr(status, 2);
    const int w = threadIdx.x >> 5;  // one record per warp: {barrier byte offset in smem, parity, block}
    status[4 + 3 * w + 0] = (int32_t)(bar - (uint32_t)__cvta_generic_to_shared(smem_dbg));
    status[4 + 3 * w + 1] = (int32_t)parity;
    status[4 + 3 * w + 2] = (int32_t)blockIdx.x;
  }
  __threadfence_system();
  // give the other warps of this CTA time to record their own stuck waits before the context dies
  for (int i = 0; i < 2000; ++i) __nanosleep(1000000);
  __trap();
}
// (try_wait with a suspend-time hint compiles to a NANOSLEEP back-off ladder: its wake-up granularity would sit on the MMA
// issuer's critical path.  The plain form is kept; the loop counts tries instead of reading the clock every iteration --
// round 1's waiting warps spent ~7 % of all issued instructions in CS2R / compare / branch around each try.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* status) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t tries = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++tries > (1u << 24)) {  // > 1 s of failed tries: from here on watch the clock and trap after ~4 s more
      const long long t0 = clock64();
      while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) mbar_timeout(bar, parity, status);
      }
      return;
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// Debug timeline of CTA 0 (one elected thread per role).  Compiled in only for the diagnostics build (-DGW_ABLATE /
// -DGW_TRACE, tools/ablate.py, tools/trace_chain.py): in the product build it costs no registers and no instructions.
#if defined(GW_ABLATE) || defined(GW_TRACE)
struct Tracer {
  long long* p;
  int n;
  __device__ __forceinline__ void init(long long* base, int role, bool on) { p = (base && on && blockIdx.x == 0) ? base + role * 2048 : nullptr, n = 0; }
  __device__ __forceinline__ void ev(int code) {
    if (p && n < 1024) {
      p[2 * n] = clock64();
      p[2 * n + 1] = code;
      ++n;
    }
  }
};
#else
struct Tracer {
  __device__ __forceinline__ void init(long long*, int, bool) {}
  __device__ __forceinline__ void ev(int) {}
};
#endif

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows are 128 B, 8-row groups are 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format [7,10)/[10,13) (0 = F16, 1 = BF16), K-major A and B,
// N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int N, int bf16) {
  uint32_t f = bf16 ? 1u : 0u;
  return (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace gw
