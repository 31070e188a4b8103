// Implicit-GEMM gate convolution, "halo" variant (sm_100a): the activation tile is loaded ONCE
// per 64-byte channel chunk, with its k//2 halo, and every tap reads it in place.
//
// The pixel tile is 8 wide x 16 tall.  A halo chunk is (8+2p) x (16+2p) pixels x 64 bytes, written
// by one 5-D TMA box (zero fill outside the image = the conv's zero padding, model.py:204-211).
// For tap (dy, dx) the UMMA A operand is the same buffer seen through a shifted descriptor: start
// address + (dy*(8+2p) + dx) * 64 bytes, 8-row groups (one tile row of 8 pixels) strided by
// SBO = (8+2p) * 64 bytes.  TMA's SWIZZLE_64B and the UMMA SW64 layout both derive the 16-byte
// chunk permutation from shared-memory address bits [7,9), so a row-shifted view stays consistent.
// L2 -> SM traffic per tile drops from taps x 8 KiB to one (8+2p)(16+2p) x 64 B halo per chunk
// (k=3: 72 -> 11.25 KiB): the dgrad conv (N = hc = 64) was bound by exactly this traffic.
//
// Optional thread-block cluster (2 or 4 CTAs, n_blocks == 1): the CTAs of a cluster walk different
// pixel tiles in lockstep and share every weight stage -- CTA r loads rows [r*n_tile/S, (r+1)*n_tile/S)
// and TMA-multicasts them to all S CTAs; a stage is recycled when the MMA warps of all S CTAs have
// released it (tcgen05.commit multicast onto every CTA's "empty" barrier).
//
// Warp roles as in nint_conv_gemm.cu (warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-11 epilogue).
#include "nint_epilogue.cuh"

namespace nint {

constexpr int kHaloCtrlBytes = 1024;
constexpr int kHaloMaxA = 12;
constexpr int kHaloMaxW = 28;

__host__ __device__ inline int halo_rows(int ksize) { return (8 + (ksize & ~1)) * (16 + (ksize & ~1)); }
static inline int halo_a_buf_bytes(const ConvGemmParams& p) {
  int m = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const int b = halo_rows(p.seg[s].ksize) * kChunkBytes;
    if (b > m) m = b;
  }
  return (m + 1023) & ~1023;
}

int conv_halo_smem_bytes(int a_buf_bytes, int na, int n_tile, int nw, int ts, int hc) {
  return 1024 + na * a_buf_bytes + nw * ts * n_tile * kChunkBytes + kHaloCtrlBytes + (4 * hc + hc) * 4;
}

// cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                                  int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <typename E, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) conv_halo_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr int CE = ElemTraits<E>::kPerChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NA = p.na_bufs, NW = p.num_stages, TS = p.taps_per_stage;
  const int w_bytes = p.n_tile * kChunkBytes;   // one tap of one chunk
  const int stage_bytes = TS * w_bytes;         // a weight stage holds up to TS consecutive taps
  uint8_t* sA = smem;
  uint8_t* sW = sA + NA * p.a_buf_bytes;
  uint8_t* ctrl = sW + NW * stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + kHaloMaxA;
  uint64_t* w_full = a_empty + kHaloMaxA;
  uint64_t* w_empty = w_full + kHaloMaxW;
  uint64_t* tfull_bar = w_empty + kHaloMaxW;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kHaloCtrlBytes - 16);
  float* s_bias = reinterpret_cast<float*>(ctrl + kHaloCtrlBytes);
  float* s_headw = s_bias + 4 * p.hc;

  const int S = p.cluster;
  const uint32_t crank = S > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = static_cast<uint16_t>((1u << S) - 1);
  // work unit of a CTA iteration = G consecutive tiles ("item group") accumulated side by side in one
  // TMEM buffer (independent accumulators: back-to-back MMAs into ONE accumulator serialise on the
  // ~190-cycle accumulate latency when N is small) and sharing every weight stage
  const int G = p.group;
  const int num_items = p.n_blocks * p.B * p.tiles_x * p.tiles_y;
  const int groups = (num_items + S * G - 1) / (S * G);
  const int clusters = gridDim.x / S;
  const int cid = blockIdx.x / S;
  const int first_item = (cid * S + static_cast<int>(crank)) * G;
  const int item_stride = clusters * S * G;
  const int items_padded = groups * S * G;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      prefetch_tensormap(&p.seg[s].tmap_act);
      prefetch_tensormap(&p.seg[s].tmap_w);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NW; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], S);   // one release per CTA of the cluster
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
  if (S > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (p.nseg > 0) {
      const bool leader = elect_one();
      int ia = 0, iw = 0;
      uint32_t pa = 0, pw = 0;
      const int w_rows = p.n_tile / S;   // rows of every weight stage this CTA fetches (and multicasts)
      for (int item = first_item; item < items_padded; item += item_stride) {
        const ItemCoord c = decode_item(p, item);   // nb is shared by the group (G > 1 only when n_blocks == 1)
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSegment& sg = p.seg[s];
          const int pad = sg.ksize >> 1;
          const int taps = sg.ksize * sg.ksize;
          const uint32_t a_bytes = static_cast<uint32_t>(halo_rows(sg.ksize) * kChunkBytes);
          int wrow = c.nb * taps * sg.nchunks * p.n_tile + static_cast<int>(crank) * w_rows;
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            mbar_wait(&a_empty[ia], pa ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&a_full[ia], a_bytes * G);
              for (int g = 0; g < G; ++g) {
                const ItemCoord cg = decode_item(p, item + g);
                tma_load_5d(sA + ia * p.a_buf_bytes + g * p.a_halo_bytes, &sg.tmap_act, &a_full[ia], ch * CE,
                            cg.x0 - pad, cg.y0 - pad, cg.b, sg.slot);
              }
            }
            if (++ia == NA) {
              ia = 0;
              pa ^= 1;
            }
            for (int tap0 = 0; tap0 < taps; tap0 += TS) {
              const int nt = (taps - tap0) < TS ? (taps - tap0) : TS;   // taps in this weight stage
              mbar_wait(&w_empty[iw], pw ^ 1);
              if (leader) mbar_arrive_expect_tx(&w_full[iw], static_cast<uint32_t>(nt * w_bytes));
              uint8_t* dst = sW + iw * stage_bytes + static_cast<int>(crank) * w_rows * kChunkBytes;
              for (int j = 0; j < nt; ++j, dst += w_bytes, wrow += p.n_tile) {
                if (leader) {
                  if (S > 1)
                    tma_load_2d_mcast(dst, &sg.tmap_w, &w_full[iw], 0, wrow, cmask);
                  else
                    tma_load_2d(dst, &sg.tmap_w, &w_full[iw], 0, wrow);
                }
              }
              if (++iw == NW) {
                iw = 0;
                pw ^= 1;
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
      int ia = 0, iw = 0;
      uint32_t pa = 0, pw = 0;
      int abuf = 0;
      uint32_t aphase = 0;
      for (int item = first_item; item < items_padded; item += item_stride) {
        mbar_wait(&tempty_bar[abuf], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(abuf * 256);
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSegment& sg = p.seg[s];
          const int ks = sg.ksize;
          const int hw = 8 + (ks & ~1);            // halo row pitch in pixels
          const uint32_t sbo = static_cast<uint32_t>(hw * kChunkBytes);
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            mbar_wait(&a_full[ia], pa);
            tc_fence_after();
            // descriptors advance by plain adds on the 16-byte-granular start-address field
            const uint64_t adesc0 = make_smem_desc_sw64(smem_u32(sA + ia * p.a_buf_bytes), 16, sbo);
            const int taps = ks * ks;
            int dy = 0, dx = 0;
            for (int tap0 = 0; tap0 < taps; tap0 += TS) {
              const int nt = (taps - tap0) < TS ? (taps - tap0) : TS;
              mbar_wait(&w_full[iw], pw);
              tc_fence_after();
              uint64_t bdesc = make_smem_desc_sw64(smem_u32(sW + iw * stage_bytes), 16, 512);
              for (int j = 0; j < nt; ++j, bdesc += static_cast<uint64_t>(w_bytes >> 4)) {
                // tap (dy, dx): the halo buffer seen through a row-shifted descriptor
                uint64_t adesc = adesc0 + static_cast<uint64_t>((dy * hw + dx) * (kChunkBytes >> 4));
                if (leader && !(p.debug_flags & 2)) {
                  for (int g = 0; g < G; ++g, adesc += static_cast<uint64_t>(p.a_halo_bytes >> 4)) {
                    umma<DT>(d_tmem + g * p.n_tile, adesc, bdesc, p.idesc, accumulate);
                    umma<DT>(d_tmem + g * p.n_tile, adesc + 2, bdesc + 2, p.idesc, 1u);
                  }
                }
                accumulate = 1;
                if (++dx == ks) {
                  dx = 0;
                  ++dy;
                }
              }
              if (leader) {
                if (S > 1)
                  umma_commit_mcast(&w_empty[iw], cmask);
                else
                  umma_commit(&w_empty[iw]);
              }
              if (++iw == NW) {
                iw = 0;
                pw ^= 1;
              }
            }
            if (leader) umma_commit(&a_empty[ia]);
            if (++ia == NA) {
              ia = 0;
              pa ^= 1;
            }
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
    conv_epilogue_loop<E, EPI>(p, warp, lane, tmem_base, tfull_bar, tempty_bar, s_bias, s_headw, first_item,
                               item_stride, items_padded, G);
  }
  tc_fence_before();
  __syncthreads();
  if (S > 1) cluster_sync_all();   // no CTA may exit while a peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// smem plan: as many weight stages as fit after NA halo buffers
void conv_halo_plan(ConvGemmParams& p) {
  p.a_halo_bytes = halo_a_buf_bytes(p);
  if (p.a_halo_bytes == 0) p.a_halo_bytes = 1024;
  // tiles per item group: fill one 256-column accumulator buffer, at most 4
  int G = 256 / p.n_tile;
  if (G > 4) G = 4;
  if (G < 1 || p.n_blocks != 1 || p.nseg == 0) G = 1;
  p.group = G;
  p.a_buf_bytes = G * p.a_halo_bytes;
  const int budget = 227 * 1024 - 1024 - kHaloCtrlBytes - (4 * p.hc + p.hc) * 4;
  const int w_bytes = p.n_tile * kChunkBytes;
  // (a) the single MMA-issuing thread pays ~300 cycles per barrier round trip: group taps so that one
  //     weight stage carries >= ~1.5k cycles of tensor work (2 MMAs of n_tile/2 cycles per tap);
  // (b) a halo chunk takes ~1.5 us to arrive (180+ scattered 64-byte rows) but only taps*n_tile cycles
  //     to consume: keep 3 weight stages and spend the rest of shared memory on halo buffers.
  int max_k = 1;
  for (int s = 0; s < p.nseg; ++s) if (p.seg[s].ksize > max_k) max_k = p.seg[s].ksize;
  int want = (1536 + p.n_tile - 1) / p.n_tile;
  if (want > max_k * max_k) want = max_k * max_k;
  int ts = 1;
  for (int cand = want; cand >= 1; --cand) {   // largest group that leaves room for 3 stages + 3 halo buffers and divides every tap count
    if (budget - 3 * cand * w_bytes < 2 * p.a_buf_bytes) continue;
    bool divides = true;
    for (int s = 0; s < p.nseg; ++s) divides = divides && ((p.seg[s].ksize * p.seg[s].ksize) % cand == 0);
    if (divides) { ts = cand; break; }
  }
  int nw = 3;
  int na = (budget - nw * ts * w_bytes) / p.a_buf_bytes;
  if (na > kHaloMaxA) na = kHaloMaxA;
  if (na < 1) na = 1;
  nw = (budget - na * p.a_buf_bytes) / (ts * w_bytes);
  if (nw > kHaloMaxW) nw = kHaloMaxW;
  p.na_bufs = na;
  p.taps_per_stage = ts;
  p.num_stages = nw;
}

template <typename E, int EPI>
static cudaError_t launch_h(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  const int smem = conv_halo_smem_bytes(p.a_buf_bytes, p.na_bufs, p.n_tile, p.num_stages, p.taps_per_stage, p.hc);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<E, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int S = p.cluster;
  const int items = p.n_blocks * p.B * p.tiles_x * p.tiles_y;
  const int groups = (items + S * p.group - 1) / (S * p.group);
  int clusters = num_sms / S;
  if (clusters > groups) clusters = groups;
  if (clusters <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * S);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_halo_kernel<E, EPI>, p);
}

cudaError_t launch_conv_halo(int epi, int dtype, const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.tile_w != 8 || p.tile_h != 16) return cudaErrorInvalidValue;
  if (p.cluster < 1 || (p.n_tile % p.cluster) || (p.cluster > 1 && p.n_blocks != 1)) return cudaErrorInvalidValue;
  if (dtype == NINT_BF16) {
    if (epi == EPI_FWD) return launch_h<__nv_bfloat16, EPI_FWD>(p, num_sms, stream);
    if (epi == EPI_BWD) return launch_h<__nv_bfloat16, EPI_BWD>(p, num_sms, stream);
    return launch_h<__nv_bfloat16, EPI_RAW>(p, num_sms, stream);
  }
  if (epi == EPI_FWD) return launch_h<float, EPI_FWD>(p, num_sms, stream);
  if (epi == EPI_BWD) return launch_h<float, EPI_BWD>(p, num_sms, stream);
  return launch_h<float, EPI_RAW>(p, num_sms, stream);
}

}  // namespace nint
