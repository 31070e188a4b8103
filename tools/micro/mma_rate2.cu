// Micro-benchmark (GPU box): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair).
#include <cstdio>
#include <cstdlib>
#include "../../nasa_niswan_b200/csrc/nint_common.cuh"
using namespace nint;

struct Cfg { int n, a_mn, b_mn, layout, nacc, sbo_a, lbo, iters, tf32, stride_a; };

__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && rank == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(NINT_BF16, 256, c.n, c.a_mn, c.b_mn);
    const uint64_t a0 = make_smem_desc(smem_u32(smem), c.lbo, c.sbo_a, c.layout);
    const uint64_t b0 = make_smem_desc(smem_u32(smem + 64 * 1024), c.lbo, 512, c.layout);
    if (leader && c.iters == 32) {
      // queue-depth probe: clock after each of 32 back-to-back MMAs issued into an idle tensor pipe
      long long ts[33];
      ts[0] = clock64();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        umma2(tm + (j % c.nacc) * c.n, a0 + ((j & 7) * c.stride_a >> 4), b0 + 2 * (j & 1), idesc, 1u);
        ts[j + 1] = clock64();
      }
      if (blockIdx.x == 0) for (int j = 0; j < 33; ++j) out[2 + j] = ts[j] - ts[0];
    }
    long long t0 = clock64();
    if (leader) {
      for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = tm + ((i + j) % c.nacc) * c.n;
          umma2(d, a0 + ((j * c.stride_a) >> 4), b0 + 2 * (j & 1), idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 0) {
    mbar_wait(&bar, 0);   // peer CTA: wait for the multicast commit
  }
  tc_fence_before(); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory"); }
}

int main() {
  long long* out; cudaMalloc(&out, 40 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
    {"2cta K-major N=256 (128/CTA) 2acc",   {256, 0, 0, 4, 2, 512, 16, 4096, 0, 32}},
    {"2cta K-major N=128 4acc",             {128, 0, 0, 4, 4, 512, 16, 4096, 0, 32}},
    {"2cta K-major N=64 4acc",              {64, 0, 0, 4, 4, 512, 16, 4096, 0, 32}},
    {"2cta K-major N=64 4acc sbo640",       {64, 0, 0, 4, 4, 640, 16, 4096, 0, 64}},
    {"2cta MN-major N=96 5acc",             {96, 1, 1, 4, 5, 512, 8192, 4096, 0, 1024}},
    {"2cta MN-major N=192 2acc",            {192, 1, 1, 4, 2, 512, 8192, 4096, 0, 1024}},
  };
  {
    Cfg q = {128, 0, 0, 4, 2, 512, 16, 32, 0, 32};
    k<<<148, 128, 200 * 1024>>>(q, out);
    cudaDeviceSynchronize();
    long long h[40];
    cudaMemcpy(h, out, 40 * 8, cudaMemcpyDeviceToHost);
    printf("issue-return clock of 32 back-to-back pair MMAs (N=128, 64 cyc each):\n  ");
    for (int j = 1; j <= 32; ++j) printf("%lld ", h[2 + j]);
    printf("\n");
    Cfg q2 = {256, 0, 0, 4, 2, 512, 16, 32, 0, 32};
    k<<<148, 128, 200 * 1024>>>(q2, out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 40 * 8, cudaMemcpyDeviceToHost);
    printf("same, N=256 (128 cyc each):\n  ");
    for (int j = 1; j <= 32; ++j) printf("%lld ", h[2 + j]);
    printf("\n");
  }
  for (auto& t : tests) {
    k<<<148, 128, 200 * 1024>>>(t.c, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-38s issue %7.1f cyc/MMA  complete %7.1f cyc/MMA  (floor %5.0f, = %5.1f per SM-tile)  %s\n", t.name,
           (double)h[0] / t.c.iters, (double)h[1] / t.c.iters, 256.0 * t.c.n / 512.0, (double)h[1] / t.c.iters / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
