// Micro-benchmark (GPU box): does tcgen05.mma.cta_group::2 issue throughput scale with the number of issuing warps?
#include <cstdio>
#include "../../nasa_niswan_b200/csrc/nint_common.cuh"
using namespace nint;

__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k(int n, int nwarps, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp < nwarps) {
    if (rank == 0) {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(NINT_BF16, 256, n, 0, 0);
      const uint64_t a0 = make_smem_desc(smem_u32(smem), 16, 512, 4);
      const uint64_t b0 = make_smem_desc(smem_u32(smem + 64 * 1024), 16, 512, 4);
      const uint32_t d = tm + warp * n;   // one accumulator per issuing warp
      long long t0 = clock64();
      if (leader) {
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) umma2(d, a0 + ((j * 32) >> 4), b0 + 2 * (j & 1), idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar[warp])), "h"((uint16_t)3) : "memory");
      }
      long long t1 = clock64();
      mbar_wait(&bar[warp], 0);
      long long t2 = clock64();
      if (leader && blockIdx.x == 0) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
    } else {
      mbar_wait(&bar[warp], 0);
    }
  }
  tc_fence_before(); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 7) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory"); }
}

int main() {
  long long* out; cudaMalloc(&out, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  for (int n : {16, 32, 64, 128, 256}) for (int nw : {1, 2, 4}) {
    if (n * nw > 512) continue;
    k<<<148, 256, 200 * 1024>>>(n, nw, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, out, 16 * 8, cudaMemcpyDeviceToHost);
    long long tmax = 0, imax = 0;
    for (int w = 0; w < nw; ++w) { if (h[2 * w + 1] > tmax) tmax = h[2 * w + 1]; if (h[2 * w] > imax) imax = h[2 * w]; }
    printf("N=%3d  %d issuing warp(s): %6.1f cyc/MMA aggregate (issue %6.1f per warp-MMA), tensor floor %d  %s\n", n, nw,
           (double)tmax / (iters * nw), (double)imax / iters, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
