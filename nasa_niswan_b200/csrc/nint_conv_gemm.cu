// Implicit-GEMM convolution on tcgen05 with fused ConvLSTM epilogues (sm_100a).
//
// Replaces, per recurrent step, the reference's cat + nn.Conv2d + split + sigmoid/tanh + state
// update chain (model.py:219-229) and, in backward, cuDNN dgrad + the autograd pointwise kernels
// (SURVEY.md section 8 a10).
//
//   D[pixel, n] = sum over segments, taps, channels  A[pixel + tap, channel] * Wp[n, tap, channel]
//
// M = 128 pixels (one tile_w x tile_h rectangle of one image), N = n_tile <= 256 accumulator
// columns in TMEM, K walks (segment, tap, 64-byte channel chunk).  A tiles come straight from the
// channels-last activation tensor through a 5-D TMA map whose out-of-bounds zero fill IS the
// convolution's zero padding; weights come from pre-packed panels through a 2-D map.  Both land
// in the 64-byte-swizzled K-major layout the UMMA descriptors expect.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over work items):
//   warp 0  TMA producer (one lane)       warp 1  MMA issuer (one lane)
//   warp 2  TMEM allocator                warps 4-11  epilogue (TMEM lane quadrant = warp % 4,
//                                                     two warps per quadrant split the channels)
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of item i
// overlaps the MMAs of item i+1.
#include "nint_epilogue.cuh"

namespace nint {

constexpr int kCtrlBytes = 1024;
constexpr int kMaxStages = 12;

__host__ __device__ inline int conv_stage_bytes(int n_tile) { return kPanelBytes + n_tile * kChunkBytes; }
__host__ __device__ inline int conv_const_bytes(int hc) { return (4 * hc + hc) * 4; }

int conv_gemm_smem_bytes(int n_tile, int num_stages, int hc) {
  return 1024 + num_stages * conv_stage_bytes(n_tile) + kCtrlBytes + conv_const_bytes(hc);
}
int conv_gemm_pick_stages(int n_tile, int hc) {
  const int budget = 227 * 1024 - 1024 - kCtrlBytes - conv_const_bytes(hc);
  int s = budget / conv_stage_bytes(n_tile);
  if (s > kMaxStages) s = kMaxStages;
  return s;
}

template <typename E, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr int CE = ElemTraits<E>::kPerChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = conv_stage_bytes(p.n_tile);
  uint8_t* ctrl = smem + p.num_stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kCtrlBytes - 16);
  float* s_bias = reinterpret_cast<float*>(ctrl + kCtrlBytes);
  float* s_headw = s_bias + 4 * p.hc;

  const int num_items = p.n_blocks * p.B * p.tiles_x * p.tiles_y;
  const uint32_t a_bytes = static_cast<uint32_t>(p.tile_w * p.tile_h * kChunkBytes);
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_tile * kChunkBytes);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      prefetch_tensormap(&p.seg[s].tmap_act);
      prefetch_tensormap(&p.seg[s].tmap_w);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (EPI == EPI_FWD) {
    for (int i = threadIdx.x; i < 4 * p.hc; i += kConvThreads) s_bias[i] = p.bias_q ? p.bias_q[i] : 0.f;
  }
  if (EPI == EPI_BWD) {
    for (int i = threadIdx.x; i < p.hc; i += kConvThreads) s_headw[i] = p.head_w ? p.head_w[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (p.nseg > 0) {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSegment& sg = p.seg[s];
          const int pad = sg.ksize >> 1;
          const int taps = sg.ksize * sg.ksize;
          int wrow = c.nb * taps * sg.nchunks * p.n_tile;
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            for (int tap = 0; tap < taps; ++tap) {
              const int dy = tap / sg.ksize - pad;
              const int dx = tap % sg.ksize - pad;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * stage_bytes;
              if (leader) {
                mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
                tma_load_5d(sa, &sg.tmap_act, &full_bar[stage], ch * CE, c.x0 + dx, c.y0 + dy, c.b, sg.slot);
                tma_load_2d(sa + kPanelBytes, &sg.tmap_w, &full_bar[stage], 0, wrow);
              }
              wrow += p.n_tile;
              if (++stage == p.num_stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (p.nseg > 0) {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int abuf = 0;
      uint32_t aphase = 0;
      int k_iters = 0;
      for (int s = 0; s < p.nseg; ++s) k_iters += p.seg[s].ksize * p.seg[s].ksize * p.seg[s].nchunks;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(&tempty_bar[abuf], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(abuf * 256);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc_sw64(sa, 16, 512);
          const uint64_t bdesc = make_smem_desc_sw64(sa + kPanelBytes, 16, 512);
          if (leader) {
            umma<DT>(d_tmem, adesc, bdesc, p.idesc, it != 0 ? 1u : 0u);
            umma<DT>(d_tmem, adesc + 2, bdesc + 2, p.idesc, 1u);
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) umma_commit(&tfull_bar[abuf]);
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    conv_epilogue_loop<E, EPI>(p, warp, lane, tmem_base, tfull_bar, tempty_bar, s_bias, s_headw, blockIdx.x,
                               gridDim.x, num_items);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <typename E, int EPI>
static cudaError_t launch_t(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  const int smem = conv_gemm_smem_bytes(p.n_tile, p.num_stages, p.hc);
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<E, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    configured = 227 * 1024;
  }
  const int items = p.n_blocks * p.B * p.tiles_x * p.tiles_y;
  const int grid = items < num_sms ? items : num_sms;
  if (grid <= 0) return cudaSuccess;
  conv_gemm_kernel<E, EPI><<<grid, kConvThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv_gemm(int epi, int dtype, const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (dtype == NINT_BF16) {
    if (epi == EPI_FWD) return launch_t<__nv_bfloat16, EPI_FWD>(p, num_sms, stream);
    if (epi == EPI_BWD) return launch_t<__nv_bfloat16, EPI_BWD>(p, num_sms, stream);
    return launch_t<__nv_bfloat16, EPI_RAW>(p, num_sms, stream);
  }
  if (epi == EPI_FWD) return launch_t<float, EPI_FWD>(p, num_sms, stream);
  if (epi == EPI_BWD) return launch_t<float, EPI_BWD>(p, num_sms, stream);
  return launch_t<float, EPI_RAW>(p, num_sms, stream);
}

}  // namespace nint
