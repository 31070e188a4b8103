// Micro-benchmark (GPU box): TMA tile-mode throughput per SM as a function of the box row size.
// One warp per CTA keeps DEPTH boxes in flight (loads) or streams stores; all 148 SMs run together.
// Tensor: channels-last [B][H][W][C] bf16 like the ConvLSTM activations; box = {row_bytes/2 ch, bw, bh}.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../../nasa_niswan_b200/csrc/nint_common.cuh"
using namespace nint;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Cfg { int iters, depth, box_bytes, tiles_x, tiles_y, B, bw, bh, store, nwarps; };

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void tma_load_2d_(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, Cfg c, long long* out, const CUtensorMap* gmap, uint8_t* gbuf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_all[128];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 128; ++i) mbar_init(&bar_all[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5;
  if (wid >= c.nwarps) return;
  uint64_t* bar = bar_all + wid * 16;
  const int box_pad = (c.box_bytes + 1023) & ~1023;
  smem += wid * c.depth * box_pad;
  const int tiles = c.tiles_x * c.tiles_y * c.B;
  long long t0 = clock64();
  const bool leader = elect_one();
  {
    if (c.store >= 3) {
      // burst of `depth` operations, then one wait.  3: 1-D bulk loads, 4: 1-D bulk stores, 5: tensor loads through a
      // descriptor in global memory, 6: 2-D tensor loads (box = {32 el, rows})
      if (c.store == 5 && leader) prefetch_tensormap(gmap);
      for (int i = 0; i < c.iters; i += c.depth) {
        if (leader && c.store != 4) mbar_arrive_expect_tx(&bar[0], c.box_bytes * c.depth);
        for (int j = 0; j < c.depth; ++j) {
          const long long t = (blockIdx.x * c.nwarps + wid + (long long)(i + j) * gridDim.x * c.nwarps) % tiles;
          if (leader) {
            if (c.store == 3) bulk_load(smem + j * box_pad, gbuf + t * c.box_bytes, c.box_bytes, &bar[0]);
            else if (c.store == 4) bulk_store(gbuf + t * c.box_bytes, smem + j * box_pad, c.box_bytes);
            else if (c.store == 5) {
              const int tx = t % c.tiles_x, ty = (t / c.tiles_x) % c.tiles_y, b = t / (c.tiles_x * c.tiles_y);
              tma_load_4d(smem + j * box_pad, gmap, &bar[0], 0, tx * 8 - 1, ty * 16 - 1, b);
            } else tma_load_2d_(smem + j * box_pad, &map, &bar[0], 0, (int)(t * c.bh));
          }
        }
        if (i == 0 && blockIdx.x == 0 && wid == 0 && leader) out[1] = clock64() - t0;
        if (c.store == 4) { if (leader) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } }
        else mbar_wait(&bar[0], (i / c.depth) & 1);
      }
      if (c.store == 4 && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (c.store == 2) {
      // burst: `depth` boxes on ONE barrier, then a single wait
      for (int i = 0; i < c.iters; i += c.depth) {
        if (leader) mbar_arrive_expect_tx(&bar[0], c.box_bytes * c.depth);
        for (int j = 0; j < c.depth; ++j) {
          const int t = (blockIdx.x * c.nwarps + wid + (i + j) * gridDim.x * c.nwarps) % tiles;
          const int tx = t % c.tiles_x, ty = (t / c.tiles_x) % c.tiles_y, b = t / (c.tiles_x * c.tiles_y);
          if (leader) tma_load_4d(smem + j * box_pad, &map, &bar[0], 0, tx * 8 - 1, ty * 16 - 1, b);
        }
        if (i == 0 && blockIdx.x == 0 && wid == 0 && leader) out[1] = clock64() - t0;
        mbar_wait(&bar[0], (i / c.depth) & 1);
      }
    } else if (!c.store) {
      for (int i = 0; i < c.iters + c.depth; ++i) {
        const int s = i % c.depth;
        if (i >= c.depth) mbar_wait(&bar[s], ((i / c.depth) - 1) & 1);
        if (i < c.iters) {
          const int t = (blockIdx.x * c.nwarps + wid + i * gridDim.x * c.nwarps) % tiles;
          const int tx = t % c.tiles_x, ty = (t / c.tiles_x) % c.tiles_y, b = t / (c.tiles_x * c.tiles_y);
          if (leader) {
            mbar_arrive_expect_tx(&bar[s], c.box_bytes);
            tma_load_4d(smem + s * box_pad, &map, &bar[s], 0, tx * 8 - 1, ty * 16 - 1, b);
          }
        }
      }
    } else {
      for (int i = 0; i < c.iters; ++i) {
        const int t = (blockIdx.x * c.nwarps + wid + i * gridDim.x * c.nwarps) % tiles;
        const int tx = t % c.tiles_x, ty = (t / c.tiles_x) % c.tiles_y, b = t / (c.tiles_x * c.tiles_y);
        if (leader) {
          tma_store_4d(&map, smem + (i % c.depth) * box_pad, 0, tx * 8, ty * 16, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (c.depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else if (c.depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
      }
      if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  __syncwarp();
  long long t1 = clock64();
  if (leader && wid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  EncodeTiledFn enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  const int H = 90, W = 144;
  long long* out;
  cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct T { const char* name; int C; int box_c; int bw, bh; int depth; int store; CUtensorMapSwizzle sw; int nwarps; int B; };
  const CUtensorMapSwizzle S32 = CU_TENSOR_MAP_SWIZZLE_32B, S64 = CU_TENSOR_MAP_SWIZZLE_64B, S128 = CU_TENSOR_MAP_SWIZZLE_128B, S0 = CU_TENSOR_MAP_SWIZZLE_NONE;
  T tests[] = {
    {"ring load 64B 8x16 d2 1w", 64, 32, 8, 16, 2, 0, S64, 1, 32},
    {"ring load 64B 8x16 d2 2w", 64, 32, 8, 16, 2, 0, S64, 2, 32},
    {"ring load 64B 8x16 d2 4w", 64, 32, 8, 16, 2, 0, S64, 4, 32},
    {"ring load 64B 8x16 d2 6w", 64, 32, 8, 16, 2, 0, S64, 6, 32},
    {"ring load 64B 8x16 d2 8w", 64, 32, 8, 16, 2, 0, S64, 8, 32},
    {"ring load 128B 8x16 d2 4w", 64, 64, 8, 16, 2, 0, S128, 4, 32},
    {"ring load 128B 8x16 d1 8w", 64, 64, 8, 16, 1, 0, S128, 8, 32},
    {"store 64B 8x16 d2 1w", 64, 32, 8, 16, 2, 1, S64, 1, 32},
    {"store 64B 8x16 d2 4w", 64, 32, 8, 16, 2, 1, S64, 4, 32},
    {"store 64B 8x16 d2 8w", 64, 32, 8, 16, 2, 1, S64, 8, 32},
    {"store 128B 8x16 d1 8w", 64, 64, 8, 16, 1, 1, S128, 8, 32},
  };
  for (auto& t : tests) {
    void* buf;
    const int B = t.B;
    const size_t bytes = (size_t)B * H * W * t.C * 2;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)t.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)t.C * 2, (cuuint64_t)W * t.C * 2, (cuuint64_t)H * W * t.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)t.box_c, (cuuint32_t)t.bw, (cuuint32_t)t.bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r;
    if (t.store == 6) {
      cuuint64_t d2[2] = {(cuuint64_t)t.box_c, (cuuint64_t)(bytes / (t.box_c * 2))};
      cuuint64_t s2[1] = {(cuuint64_t)t.box_c * 2};
      cuuint32_t b2[2] = {(cuuint32_t)t.box_c, (cuuint32_t)t.bh};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, d2, s2, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, t.sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, t.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", t.name, (int)r); continue; }
    Cfg c;
    c.iters = 512; c.depth = t.depth; c.box_bytes = t.box_c * 2 * t.bw * t.bh; c.tiles_x = W / 8; c.tiles_y = (H + 15) / 16; c.B = B;
    c.bw = t.bw; c.bh = t.bh; c.store = t.store; c.nwarps = t.nwarps;
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (t.store >= 3 && t.store != 5) { c.tiles_x = 1; c.tiles_y = 1; c.B = (int)(bytes / c.box_bytes); if (t.store == 6) c.bh = t.bh; }
    for (int rep = 0; rep < 2; ++rep) k<<<148, 256, 200 * 1024>>>(m, c, out, gm, (uint8_t*)buf);
    cudaError_t e = cudaDeviceSynchronize();
    long long hh[2] = {0, 0};
    cudaMemcpy(hh, out, 16, cudaMemcpyDeviceToHost);
    const long long h = hh[0];
    const double cyc = (double)h / c.iters;
    const int rows = t.bw * t.bh;
    printf("%-46s %8.0f cyc/box/warp  %6.2f cyc/row(SM)  %6.1f B/cyc/SM  %s\n", t.name, cyc, cyc / rows / t.nwarps, t.nwarps * c.box_bytes / cyc,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (t.store == 2) printf("      first burst issued after %lld cycles (%d boxes)\n", hh[1], t.depth);
    cudaFree(buf);
  }
  return 0;
}
