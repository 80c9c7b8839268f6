// tma.cuh — minimal mbarrier + 1-D bulk-copy (TMA) wrappers for sm_100a.
// A warp streams the per-timestep slabs of its 32 trajectories (contiguous in the
// [k][slot][component] layout) into a shared-memory ring several steps ahead of use:
// one elected lane arms an mbarrier with the byte count and issues cp.async.bulk
// (SASS: UBLKCP); all lanes wait on the barrier's phase parity before reading.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ilqr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global → shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Ring-stage reuse.  The bulk copy that refills a stage (async proxy) must not be issued before the shared-memory reads of
// that stage have RETURNED.  __syncwarp() only orders the instruction streams: an LDS is complete when its wavefronts
// have come back through the load/store pipe, and when other warps on the SM keep that pipe busy with uncoalesced
// global accesses (the gather kernels of the batch path running in another stream; the slot copies of the round kernel)
// a few 4-lane wavefronts can return after a refill that hit in L2 has landed — those lanes then see time step k + D
// instead of k.  Found with tools/trace_divergence.py (concurrent handles deviated from a solo solve in groups of four
// adjacent slots, from some time step to the end of the horizon).  Fix: the elected lane consumes one register of every
// LDS of the stage (a warp instruction retires for all lanes at once) in a store the compiler cannot drop, then issues.
template <int C> __device__ __forceinline__ uint32_t ring_token(const double* v) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < C; i += 2) t |= (uint32_t)__double2hiint(v[i]);   // one element per 16-byte LDS
  return t;
}
__device__ __forceinline__ void ring_reads_done(volatile uint32_t* scratch, uint32_t token) { *scratch = token; }

}  // namespace ilqr
