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
#include "nint_common.cuh"
#include "nint_kernels.h"

namespace nint {

constexpr int kConvThreads = 384;
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

struct ItemCoord {
  int nb, b, x0, y0;
};
__device__ __forceinline__ ItemCoord decode_item(const ConvGemmParams& p, int item) {
  ItemCoord c;
  c.nb = item % p.n_blocks;
  int r = item / p.n_blocks;
  const int tx = r % p.tiles_x;
  r /= p.tiles_x;
  const int ty = r % p.tiles_y;
  c.b = r / p.tiles_y;
  c.x0 = tx * p.tile_w;
  c.y0 = ty * p.tile_h;
  return c;
}

template <typename E, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr int CE = ElemTraits<E>::kPerChunk;
  constexpr bool FAST = (DT == NINT_BF16);
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
    if (lane == 0 && p.nseg > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSegment& sg = p.seg[s];
          const int pad = sg.ksize >> 1;
          const int taps = sg.ksize * sg.ksize;
          int wrow = c.nb * taps * sg.nchunks * p.n_tile;
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = tap / sg.ksize - pad;
            const int dx = tap % sg.ksize - pad;
            for (int ch = 0; ch < sg.nchunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * stage_bytes;
              mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
              tma_load_5d(sa, &sg.tmap_act, &full_bar[stage], ch * CE, c.x0 + dx, c.y0 + dy, c.b, sg.slot);
              tma_load_2d(sa + kPanelBytes, &sg.tmap_w, &full_bar[stage], 0, wrow);
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
    if (lane == 0 && p.nseg > 0) {
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
          const uint32_t sb = sa + kPanelBytes;
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const uint64_t adesc = make_smem_desc_sw64(sa + k2 * 32, 16, 512);
            const uint64_t bdesc = make_smem_desc_sw64(sb + k2 * 32, 16, 512);
            umma<DT>(d_tmem, adesc, bdesc, p.idesc, (it | k2) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[abuf]);
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the
    // 16-channel groups between them (half = 0 / 1).
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int ty = row / p.tile_w;
    const int tx = row - ty * p.tile_w;
    const int hc = p.hc;
    const int hcb = p.hcb;
    int abuf = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      const int y = c.y0 + ty, x = c.x0 + tx;
      const bool valid = (ty < p.tile_h) && (y < p.H) && (x < p.W);
      const long long pix = (static_cast<long long>(c.b) * p.H + y) * p.W + x;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(abuf * 256);
      bool waited = (p.nseg == 0);
      auto wait_acc = [&]() {
        if (!waited) {
          mbar_wait(&tfull_bar[abuf], aphase);
          tc_fence_after();
          waited = true;
        }
      };
      if constexpr (EPI == EPI_FWD) {
        // model.py:221-229.  columns of this n-block: gate * hcb + cc
        const float* cprev = p.c_prev ? p.c_prev + pix * hc + c.nb * hcb : nullptr;
        float* cout = p.c_out + pix * hc + c.nb * hcb;
        E* hout = reinterpret_cast<E*>(p.h_out) + pix * p.hc_pad + c.nb * hcb;
        E* gout = p.gates_out ? reinterpret_cast<E*>(p.gates_out) + pix * 4 * hc + c.nb * p.n_tile : nullptr;
        const float* bq = s_bias + c.nb * p.n_tile;
        for (int cg = half * 16; cg < hcb; cg += 32) {
          float cn[16];
          if (cprev && valid) {   // issued before the accumulator wait: overlaps the MMA tail
            load_elems<float, 16>(cprev + cg, cn);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) cn[j] = 0.f;
          }
          wait_acc();
          float a[4][16];
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld16(taddr + g * hcb + cg, a[g]);
          tmem_ld_wait();
          if (valid) {
            float hn[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float gi = act_sigmoid<FAST>(a[0][j] + bq[cg + j]);
              const float gf = act_sigmoid<FAST>(a[1][j] + bq[hcb + cg + j]);
              const float gg = act_tanh<FAST>(a[2][j] + bq[2 * hcb + cg + j]);
              const float go = act_sigmoid<FAST>(a[3][j] + bq[3 * hcb + cg + j]);
              const float cv = fmaf(cn[j], gf, gi * gg);
              cn[j] = cv;
              float hv = go * act_tanh<FAST>(cv);
              if constexpr (DT == NINT_TF32) hv = round_tf32(hv);  // h feeds the next step's tf32 MMA
              hn[j] = hv;
              a[0][j] = gi; a[1][j] = gf; a[2][j] = gg; a[3][j] = go;
            }
            store_elems<float, 16>(cout + cg, cn);
            store_elems<E, 16>(hout + cg, hn);
            if (gout) {
#pragma unroll
              for (int g = 0; g < 4; ++g) store_elems<E, 16>(gout + g * hcb + cg, a[g]);
            }
          }
        }
        wait_acc();   // warps without a channel group (hcb == 16) still take part in the handshake
      } else if constexpr (EPI == EPI_BWD) {
        // gate backward; accumulator column c = dh_t[c] from the dgrad conv
        const E* gin = reinterpret_cast<const E*>(p.gates_in) + pix * 4 * hc;
        E* dgo = reinterpret_cast<E*>(p.dgates_out) + pix * 4 * hc;
        const float* ccur = p.c_cur + pix * hc;
        const float* cprv = p.c_prev_b ? p.c_prev_b + pix * hc : nullptr;
        const float* dcin = p.dc_in ? p.dc_in + pix * hc : nullptr;
        float* dcout = p.dc_out + pix * hc;
        float dpred = 0.f;
        if (p.head_dpred && valid) {
          const long long hw = static_cast<long long>(p.H) * p.W;
          dpred = p.head_dpred[c.b * p.head_dpred_bstride + (pix - c.b * hw)];
        }
        for (int c0 = half * 16; c0 < hc; c0 += 32) {
          const int nb = c0 / hcb, cc = c0 - nb * hcb;
          const int qb = nb * 4 * hcb + cc;  // + gate * hcb
          float gi[16], gf[16], gg[16], go[16], ct[16], cp[16], dc[16], dh[16];
          if (valid) {   // all global loads of the group in flight before the accumulator wait
            load_elems<E, 16>(gin + qb, gi);
            load_elems<E, 16>(gin + qb + hcb, gf);
            load_elems<E, 16>(gin + qb + 2 * hcb, gg);
            load_elems<E, 16>(gin + qb + 3 * hcb, go);
            load_elems<float, 16>(ccur + c0, ct);
            if (cprv) {
              load_elems<float, 16>(cprv + c0, cp);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) cp[j] = 0.f;
            }
            if (dcin) {
              load_elems<float, 16>(dcin + c0, dc);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dc[j] = 0.f;
            }
          }
          if (p.nseg > 0) {
            wait_acc();
            tmem_ld16(taddr + c0, dh);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = 0.f;
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float dhv = fmaf(dpred, s_headw[c0 + j], dh[j]);
              const float tc = act_tanh<FAST>(ct[j]);
              const float d_o = dhv * tc;
              const float dcv = fmaf(dhv * go[j], 1.f - tc * tc, dc[j]);
              const float d_i = dcv * gg[j];
              const float d_g = dcv * gi[j];
              const float d_f = dcv * cp[j];
              dc[j] = dcv * gf[j];
              const float i_ = gi[j], f_ = gf[j], g_ = gg[j], o_ = go[j];
              gi[j] = d_i * i_ * (1.f - i_);
              gf[j] = d_f * f_ * (1.f - f_);
              gg[j] = d_g * (1.f - g_ * g_);
              go[j] = d_o * o_ * (1.f - o_);
              if constexpr (DT == NINT_TF32) {
                // dgates are MMA operands of dgrad and wgrad: round to nearest tf32 (the MMA truncates)
                gi[j] = round_tf32(gi[j]); gf[j] = round_tf32(gf[j]);
                gg[j] = round_tf32(gg[j]); go[j] = round_tf32(go[j]);
              }
            }
            store_elems<float, 16>(dcout + c0, dc);
            store_elems<E, 16>(dgo + qb, gi);
            store_elems<E, 16>(dgo + qb + hcb, gf);
            store_elems<E, 16>(dgo + qb + 2 * hcb, gg);
            store_elems<E, 16>(dgo + qb + 3 * hcb, go);
          }
        }
        wait_acc();
      } else {
        float* ro = p.raw_out + pix * (p.n_blocks * p.n_tile) + c.nb * p.n_tile;
        wait_acc();
        for (int c0 = half * 16; c0 < p.n_tile; c0 += 32) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (valid) store_elems<float, 16>(ro + c0, v);
        }
      }
      if (p.nseg > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[abuf]);
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
      }
    }
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
