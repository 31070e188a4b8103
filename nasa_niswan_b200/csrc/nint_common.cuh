// Shared device-side primitives for the sm_100a kernels: mbarrier, TMA (tiled tensor
// loads + 1-D bulk copies), tcgen05 (TMEM alloc, UMMA issue/commit, TMEM loads) and
// the shared-memory / instruction descriptor encodings.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace nint {

// ---------------------------------------------------------------------------------
// constants of the operand layout shared by every tensor-core kernel in this repo:
// one K (or MN) "chunk" is 64 bytes of channels of one pixel (32 bf16 / 16 tf32),
// staged in shared memory with the 64-byte swizzle, 128 pixels (8 KiB) per panel.
// ---------------------------------------------------------------------------------
constexpr int kChunkBytes = 64;
constexpr int kTilePixels = 128;                       // UMMA M (conv kernels) / K-tile (wgrad)
constexpr int kPanelBytes = kTilePixels * kChunkBytes; // 8 KiB
constexpr int kTmemCols = 512;

enum : int { NINT_BF16 = 0, NINT_TF32 = 1 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  Role loops run warp-uniform (all 32 lanes walk the same
// control flow, so addresses / descriptors live in uniform registers) and only the instructions with
// side effects are predicated on the elected lane; guarding the whole loop with `lane == 0` instead
// costs ~100 cycles per tcgen05.mma in R2UR moves and dependent scalar chains (ncu, profiles/).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// suspend-time hint: a waiting thread may sleep up to this long inside one try_wait (it is woken when the phase
// completes), instead of re-polling: 14 of a CTA's 16 warps are waiting at any time and every poll costs issue slots
// and power on a part that runs power-capped
constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
// The message of a timed-out wait is a device printf, i.e. a CALL inside a function that is inlined into every hot loop of
// every kernel; it exists in the experiment build only (-DNINT_KNOBS=1), the product build just traps.
#if defined(NINT_KNOBS) && NINT_KNOBS
#define NINT_TIMEOUT_PRINTF(...) printf(__VA_ARGS__)
#else
#define NINT_TIMEOUT_PRINTF(...) ((void)0)
#endif
// Bounded wait: a protocol bug must trap (error returned to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      NINT_TIMEOUT_PRINTF("nint: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
                          (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// Bounded spin on a shared-memory progress counter (in-order hand-off between two MMA-issuing warps): like
// mbar_wait, a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void spin_until_at_least(volatile uint32_t* counter, uint32_t value) {
  if (*counter >= value) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (*counter < value) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      NINT_TIMEOUT_PRINTF("nint: progress counter wait timed out (block %d thread %d want %u have %u)\n", (int)blockIdx.x,
                          (int)threadIdx.x, value, *counter);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 5-D tiled load, coordinates innermost first: (channel, x, y, image, slot)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// 2-D tiled load (packed weight panels): coordinates (element in chunk, row)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 3-D tiled load (packed weight panels): coordinates (element in chunk, row of the N tile, chunk-tap index)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem];  kind::f16 covers bf16/fp16 inputs, kind::tf32 fp32-stored tf32.
template <int DTYPE>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  if constexpr (DTYPE == NINT_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------
// descriptors (bit layouts: CUTLASS cute/arch/mma_sm100_desc.hpp, restated)
// ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 64-byte swizzle.
//   [0,14)  start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4     [46,48) version = 1     [61,64) layout (4 = SWIZZLE_64B)
// K-major operand (conv kernels):  rows are 64 B, 8-row groups SBO apart (dense: 512 B), LBO = 16 B
//                                  (ignored by the hardware for swizzled K-major; CUTLASS encodes 1).
// MN-major operand (wgrad):        64-byte (one chunk) MN atoms LBO apart (one panel), 8-k groups SBO
//                                  apart (512 B).
// layout codes: 4 = SWIZZLE_64B, 1 = SWIZZLE_128B_BASE32B (128-byte rows swizzled in 32-byte units: the
// only MN-major layout the hardware accepts for tf32 operands; TMA twin: SWIZZLE_128B_ATOM_32B).
constexpr uint32_t kLayoutSw64 = 4, kLayoutSw128Base32 = 1;
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
__host__ __device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr, uint32_t lbo_bytes,
                                                                 uint32_t sbo_bytes) {
  return make_smem_desc(smem_addr, lbo_bytes, sbo_bytes, kLayoutSw64);
}
// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (bf16 = 1, tf32 = 2)
//   [15] A major  [16] B major (0 = K-major, 1 = MN-major)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t make_idesc(int dtype, int m, int n, int a_mn_major, int b_mn_major) {
  const uint32_t fmt = (dtype == NINT_BF16) ? 1u : 2u;
  uint32_t d = 0;
  d |= 1u << 4;
  d |= fmt << 7;
  d |= fmt << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

// 16-byte sub-chunk permutation of the 64-byte swizzle for a row at byte offset `row_off`
// (row_off multiple of 64) inside a 512-byte aligned tile: chunk j lives at j ^ ((row_off >> 7) & 3).
__host__ __device__ __forceinline__ int sw64_chunk(int row, int j) { return j ^ ((row >> 1) & 3); }

// ---------------------------------------------------------------------------------
// element helpers (E = __nv_bfloat16 or float)
// ---------------------------------------------------------------------------------
template <typename E>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr int kDtype = NINT_BF16;
  static constexpr int kPerChunk = 32;
};
template <>
struct ElemTraits<float> {
  static constexpr int kDtype = NINT_TF32;
  static constexpr int kPerChunk = 16;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}

__device__ __forceinline__ void lds128(uint32_t saddr, float* v) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(saddr));
}
// 256-bit global accesses (sm_100: LDG/STG.256), 32-byte aligned
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// store / load N consecutive elements as fp32 registers <-> E in global memory.  N*sizeof(E) must be a
// multiple of 16 bytes; multiples of 32 bytes (32-byte aligned addresses) use 256-bit accesses.
template <typename E, int N>
__device__ __forceinline__ void store_elems(E* dst, const float* v) {
  constexpr int BYTES = N * sizeof(E);
  if constexpr (sizeof(E) == 2) {
    if constexpr (BYTES % 32 == 0) {
#pragma unroll
      for (int i = 0; i < N; i += 16) {
        uint32_t u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(v[i + 2 * j], v[i + 2 * j + 1]);
        stg256(dst + i, u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[i + 0], v[i + 1]);
        u.y = pack_bf16x2(v[i + 2], v[i + 3]);
        u.z = pack_bf16x2(v[i + 4], v[i + 5]);
        u.w = pack_bf16x2(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  } else {
    if constexpr (BYTES % 32 == 0) {
#pragma unroll
      for (int i = 0; i < N; i += 8) stg256(dst + i, reinterpret_cast<const uint32_t*>(v + i));
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
}
template <typename E, int N>
__device__ __forceinline__ void load_elems(const E* src, float* v) {
  constexpr int BYTES = N * sizeof(E);
  if constexpr (sizeof(E) == 2) {
    if constexpr (BYTES % 32 == 0) {
#pragma unroll
      for (int i = 0; i < N; i += 16) {
        uint32_t u[8];
        ldg256(src + i, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) unpack_bf16x2(u[j], v[i + 2 * j], v[i + 2 * j + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + i);
        unpack_bf16x2(u.x, v[i + 0], v[i + 1]);
        unpack_bf16x2(u.y, v[i + 2], v[i + 3]);
        unpack_bf16x2(u.z, v[i + 4], v[i + 5]);
        unpack_bf16x2(u.w, v[i + 6], v[i + 7]);
      }
    }
  } else {
    if constexpr (BYTES % 32 == 0) {
#pragma unroll
      for (int i = 0; i < N; i += 8) ldg256(src + i, reinterpret_cast<uint32_t*>(v + i));
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        const float4 f = *reinterpret_cast<const float4*>(src + i);
        v[i] = f.x; v[i + 1] = f.y; v[i + 2] = f.z; v[i + 3] = f.w;
      }
    }
  }
}

// round-to-nearest tf32 (tcgen05 kind::tf32 truncates the low 13 mantissa bits of its operands)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// activations.  FAST: one MUFU.TANH each (bf16 mode, error below bf16 rounding).
// precise: ex2 + rcp, ~1e-6 absolute (fp32/tf32 mode).
template <bool FAST>
__device__ __forceinline__ float act_tanh(float x) {
  if constexpr (FAST) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    const float e = __expf(-2.0f * fabsf(x));          // in (0,1]
    const float r = __fdividef(1.0f - e, 1.0f + e);
    return copysignf(r, x);
  }
}
// sigmoid of a pre-activation that arrives already HALVED: the packed forward weights and biases of the i, f, o gates
// are scaled by 0.5 (exact: a power of two commutes with every rounding), so sigma(a) = 0.5 * tanh(a / 2) + 0.5 needs no
// multiply in the epilogue -- one MUFU + one FFMA per gate value
template <bool FAST>
__device__ __forceinline__ float act_sigmoid_halved(float half_x) {
  if constexpr (FAST) {
    return fmaf(act_tanh<true>(half_x), 0.5f, 0.5f);
  } else {
    return __fdividef(1.0f, 1.0f + __expf(-2.0f * half_x));
  }
}
template <bool FAST>
__device__ __forceinline__ float act_sigmoid(float x) {
  if constexpr (FAST) {
    return fmaf(act_tanh<true>(0.5f * x), 0.5f, 0.5f);
  } else {
    return __fdividef(1.0f, 1.0f + __expf(-x));
  }
}

}  // namespace nint
