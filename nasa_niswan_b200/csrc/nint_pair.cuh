// CTA-pair (thread-block cluster of 2, tcgen05 cta_group::2) primitives.  sm_100a only.
//
// A pair MMA has M = 256: rows [0,128) come from the A tile in CTA 0's shared memory, rows [128,256)
// from the same offset in CTA 1's; each CTA holds HALF of the B operand (N/2 rows of N) and receives its
// own 128 accumulator rows x N columns in its own TMEM.  Only the leader (cluster rank 0) issues MMAs and
// commits; operand "full" barriers live in the leader and count the bytes of both CTAs' TMA loads
// (measured on B200, tools/micro/mma_rate2.cu: the pair MMA runs at the tensor pipe's nominal rate where
// a single-CTA M=128 MMA pays ~45-60 extra cycles per instruction).
#pragma once
#include "nint_common.cuh"

namespace nint {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive on an mbarrier anywhere in the cluster (shared::cluster address).  Relaxed: the only data the waiter
// (the MMA issuer) depends on are this warp's tcgen05.ld reads of the accumulator, ordered by the preceding
// tcgen05.fence::before_thread_sync; a .release.cluster arrive costs a MEMBAR + ERRBAR per call (ncu: ~15 % of
// the epilogue's stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMEM allocation for the pair: one warp of EACH CTA executes these
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMA loads whose completion bytes are counted on a barrier given by its shared::cluster address (the
// leader's): data lands in the executing CTA's own shared memory.
__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                 int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                 int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// D[tmem of both CTAs] (+)= A[256 x K: 128 rows per CTA] * B[N x K: N/2 rows per CTA]; leader CTA only
template <int DTYPE>
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (DTYPE == NINT_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// all pair MMAs issued so far by this thread arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// tcgen05.mma with the two shared-memory descriptors passed as (low, high) 32-bit halves: along a K walk only the
// low words (start-address field) change, so the issuing thread's per-MMA work is two 32-bit adds and two
// register->uniform moves instead of 64-bit descriptor arithmetic (the issue loop must run faster than the
// 64-cycle N=128 pair MMA or every barrier round trip shows up as tensor idle time).
template <int DTYPE, bool PAIR>
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                          uint32_t idesc, uint32_t accumulate) {
#define NINT_UMMA_LOHI(GROUP, KIND)                                                                         \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %6, 0;\n\t"                       \
               "mov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"                                        \
               "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], ad, bd, %5, p;\n\t}" ::"r"(d_tmem),   \
               "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)                          \
               : "memory")
  if constexpr (DTYPE == NINT_BF16) {
    if constexpr (PAIR) NINT_UMMA_LOHI("2", "f16"); else NINT_UMMA_LOHI("1", "f16");
  } else {
    if constexpr (PAIR) NINT_UMMA_LOHI("2", "tf32"); else NINT_UMMA_LOHI("1", "tf32");
  }
#undef NINT_UMMA_LOHI
}

// ---- "elected" variants: every lane of the (converged) warp executes the call, the instruction itself is
// predicated on elect.sync inside the asm block.  The surrounding C++ then has no divergent branch, which lets
// ptxas keep descriptors / addresses in uniform registers (a lane-predicated branch around tcgen05.mma cost ~8
// R2UR moves and ~60 extra cycles per tap: timeline traces, tools/trace_report.py).
template <int DTYPE, bool PAIR>
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  if constexpr (DTYPE == NINT_BF16) {
    if constexpr (PAIR)
      asm volatile(
          "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
  } else {
    if constexpr (PAIR)
      asm volatile(
          "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "@e tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
  }
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  if constexpr (PAIR)
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}

}  // namespace nint
