// Micro-benchmark (GPU box): issue rate of tcgen05.mma for the shapes the ConvLSTM kernels use.
// One CTA per SM (148), garbage operands in shared memory, `iters` MMAs issued back to back by one
// elected thread, then one commit; reports cycles per MMA.   nvcc -arch=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../../nasa_niswan_b200/csrc/nint_common.cuh"
using namespace nint;

struct Cfg { int n, a_mn, b_mn, layout, nacc, sbo_a, lbo, iters, tf32, stride_a; };

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(c.tf32 ? NINT_TF32 : NINT_BF16, 128, c.n, c.a_mn, c.b_mn);
    const uint64_t a0 = make_smem_desc(smem_u32(smem), c.lbo, c.sbo_a, c.layout);
    const uint64_t b0 = make_smem_desc(smem_u32(smem + 64 * 1024), c.lbo, 512, c.layout);
    long long t0 = clock64();
    if (leader) {
      for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = tm + ((i + j) % c.nacc) * c.n;
          const uint64_t a = a0 + ((j * c.stride_a) >> 4);
          if (c.tf32) umma<NINT_TF32>(d, a, b0 + 2 * (j & 1), idesc, 1u);
          else umma<NINT_BF16>(d, a, b0 + 2 * (j & 1), idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
    {"K-major sw64 N=256 1acc",            {256, 0, 0, 4, 1, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=256 2acc",            {256, 0, 0, 4, 2, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=128 1acc",            {128, 0, 0, 4, 1, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=64  1acc",            {64, 0, 0, 4, 1, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=64  4acc",            {64, 0, 0, 4, 4, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=64  8acc",            {64, 0, 0, 4, 8, 512, 16, 4096, 0, 32}},
    {"K-major sw64 N=64  4acc sbo640",     {64, 0, 0, 4, 4, 640, 16, 4096, 0, 64}},
    {"K-major sw64 N=256 2acc sbo640",     {256, 0, 0, 4, 2, 640, 16, 4096, 0, 64}},
    {"MN-major sw64 N=96 1acc",            {96, 1, 1, 4, 1, 512, 8192, 4096, 0, 1024}},
    {"MN-major sw64 N=96 5acc",            {96, 1, 1, 4, 5, 512, 8192, 4096, 0, 1024}},
    {"MN-major sw64 N=96 5acc sbo640",     {96, 1, 1, 4, 5, 640, 12288, 4096, 0, 1280}},
    {"MN-major sw64 N=256 2acc",           {256, 1, 1, 4, 2, 512, 8192, 4096, 0, 1024}},
    {"A K-major, B MN-major N=96 5acc",    {96, 0, 1, 4, 5, 512, 8192, 4096, 0, 32}},
    {"A MN-major, B K-major N=96 5acc",    {96, 1, 0, 4, 5, 512, 8192, 4096, 0, 1024}},
    {"tf32 K-major sw64 N=256 2acc",       {256, 0, 0, 4, 2, 512, 16, 4096, 1, 32}},
    {"tf32 MN-major sw128b32 N=96 5acc",   {96, 1, 1, 1, 5, 512, 16384, 4096, 1, 1024}},
  };
  for (auto& t : tests) {
    k<<<148, 128, 200 * 1024>>>(t.c, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double ideal = 128.0 * t.c.n / 256.0 * (t.c.tf32 ? 1.0 : 1.0);
    printf("%-38s issue %7.1f cyc/MMA  complete %7.1f cyc/MMA  (tensor floor %5.0f)  %s\n", t.name,
           (double)h[0] / t.c.iters, (double)h[1] / t.c.iters, ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
