// Implicit-GEMM gate convolution, "halo" variant (sm_100a): the activation tile is loaded ONCE
// per 64-byte channel chunk, with its k//2 halo, and every tap reads it in place.
//
// The pixel tile is 8 wide x 16 tall.  A halo chunk is (8+2p) x (16+2p) pixels x 64 bytes, written
// by one 5-D TMA box (zero fill outside the image = the conv's zero padding, model.py:204-211).
// For tap (dy, dx) the UMMA A operand is the same buffer seen through a shifted descriptor: start
// address + (dy*(8+2p) + dx) * 64 bytes, 8-row groups (one tile row of 8 pixels) strided by
// SBO = (8+2p) * 64 bytes.  TMA's SWIZZLE_64B and the UMMA SW64 layout both derive the 16-byte
// chunk permutation from shared-memory address bits [7,9) (verified on hardware: the descriptor
// base-offset field must stay 0), so a row-shifted view stays consistent.
//
// CTA pair (cluster of 2, tcgen05 cta_group::2): measured on B200 (tools/micro/mma_rate*.cu) a
// single-CTA M=128 MMA costs 171 / 110 / 88 cycles at N = 256 / 128 / 64 -- a ~60-cycle fixed cost
// per instruction -- while the pair MMA (M=256: 128 pixels from each CTA) runs at the nominal floor
// (128 / 64 / 43 cycles for BOTH SMs).  In pair mode the two CTAs walk different pixel tiles in
// lockstep; each loads its own halo tiles and HALF of every weight stage (rows [r*N/2, (r+1)*N/2)),
// which also halves the weight bytes every SM has to ingest.  The leader CTA (rank 0) issues the MMAs;
// "full" barriers live in the leader (both CTAs' TMA loads complete_tx on them), "empty"/accumulator
// barriers are signalled in both CTAs by multicast tcgen05.commit.
//
// Item groups: G consecutive tiles (G * n_tile <= 256 columns) are accumulated side by side in one
// TMEM buffer and share every weight stage.
//
// Warp roles: warp 0 operand TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator + epilogue TMA stores,
// warp 3 epilogue TMA loads, warps 4-11 epilogue math (nint_epilogue.cuh).
#include "nint_epilogue.cuh"

namespace nint {

constexpr int kHaloCtrlBytes = 1024;
constexpr int kHaloMaxA = 12;
constexpr int kHaloMaxW = 28;
constexpr int kMaxAcc = 4;       // TMEM accumulator buffers (512 columns / acc_cols, at most 4)

__host__ __device__ inline int halo_rows(int ksize) { return (8 + (ksize & ~1)) * (16 + (ksize & ~1)); }
static inline int halo_a_buf_bytes(const ConvGemmParams& p) {
  int m = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const int b = halo_rows(p.seg[s].ksize) * kChunkBytes;
    if (b > m) m = b;
  }
  return (m + 1023) & ~1023;
}

static int halo_operand_bytes(int a_buf_bytes, int na, int n_tile, int nw, int ts, int cluster) {
  return (na * a_buf_bytes + nw * ts * (n_tile / cluster) * kChunkBytes + 1023) & ~1023;
}
int conv_halo_smem_bytes(const ConvGemmParams& p) {
  return 1024 + halo_operand_bytes(p.a_buf_bytes, p.na_bufs, p.n_tile, p.num_stages, p.taps_per_stage, p.cluster) +
         p.e_stages * p.e_stage_bytes + kHaloCtrlBytes + (4 * p.hc + p.hc) * 4;
}

__device__ __forceinline__ int halo_operand_bytes_dev(int a_buf_bytes, int na, int stage_bytes, int nw) {
  return (na * a_buf_bytes + nw * stage_bytes + 1023) & ~1023;
}

template <typename E, int EPI, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1) conv_halo_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr int CE = ElemTraits<E>::kPerChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int S = PAIR ? 2 : 1;               // 1: single CTA, 2: CTA pair (cta_group::2)
  constexpr bool pair = PAIR;
  const int NA = p.na_bufs, NW = p.num_stages, TS = p.taps_per_stage;
  const int w_rows = p.n_tile / S;              // weight rows (of N) this CTA holds
  const int w_bytes = w_rows * kChunkBytes;     // one tap of one chunk, this CTA's share
  const int stage_bytes = TS * w_bytes;         // a weight stage holds up to TS consecutive taps
  uint8_t* sA = smem;
  uint8_t* sW = sA + NA * p.a_buf_bytes;
  uint8_t* sE = smem + halo_operand_bytes_dev(p.a_buf_bytes, NA, stage_bytes, NW);   // epilogue I/O stages
  uint8_t* ctrl = sE + p.e_stages * p.e_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + kHaloMaxA;
  uint64_t* w_full = a_empty + kHaloMaxA;
  uint64_t* w_empty = w_full + kHaloMaxW;
  uint64_t* tfull_bar = w_empty + kHaloMaxW;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint64_t* e_full = tempty_bar + kMaxAcc;
  uint64_t* e_empty = e_full + kEpiMaxStages;
  uint64_t* st_ready = e_empty + kEpiMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kHaloCtrlBytes - 16);
  float* s_bias = reinterpret_cast<float*>(ctrl + kHaloCtrlBytes);
  float* s_headw = s_bias + 4 * p.hc;

  uint32_t crank = 0;
  if constexpr (pair) crank = cluster_ctarank();
  const bool lead_cta = crank == 0;
  // work unit of a CTA iteration = G consecutive tiles ("tile group") accumulated side by side in one
  // TMEM buffer and sharing every weight stage; the two CTAs of a pair take adjacent groups
  const int G = p.group;
  const TileWalk walk = make_walk(p, S, static_cast<int>(crank));
  const bool resident = p.w_resident != 0;   // the whole weight slice of this n-block fits: load it once

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      prefetch_tensormap(&p.seg[s].tmap_act);
      prefetch_tensormap(&p.seg[s].tmap_w);
    }
  }
  if (warp == 3 && lane == 0 && EPI != EPI_RAW) {
    prefetch_tensormap(&p.tm_c);
    prefetch_tensormap(&p.tm_g);
    prefetch_tensormap(EPI == EPI_FWD ? &p.tm_h : &p.tm_dc);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NW; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8 * S);   // the leader waits for the epilogue warps of both CTAs
    }
    for (int s = 0; s < kEpiMaxStages; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 1);
      mbar_init(&st_ready[s], 8);         // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (pair) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  if (EPI == EPI_FWD) {
    for (int i = threadIdx.x; i < 4 * p.hc; i += kConvThreads) s_bias[i] = p.bias_q ? p.bias_q[i] : 0.f;
  }
  if (EPI == EPI_BWD) {
    for (int i = threadIdx.x; i < p.hc; i += kConvThreads) s_headw[i] = p.head_w ? p.head_w[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ operand TMA producer (both CTAs)
    if (p.nseg > 0) {
      const bool leader = elect_one();
      int ia = 0, iw = 0;
      uint32_t pa = 0, pw = 0;
      int es = 0;          // epilogue stage ring (forward: this warp also prefetches c_{t-1})
      uint32_t eph = 0;
      bool first = true;
      for (int base = walk.first_tile; base < walk.tiles_padded; base += walk.tile_stride) {
        for (int s = 0; s < p.nseg; ++s) {
          const ConvSegment& sg = p.seg[s];
          const int pad = sg.ksize >> 1;
          const int taps = sg.ksize * sg.ksize;
          const uint32_t a_bytes = static_cast<uint32_t>(halo_rows(sg.ksize) * kChunkBytes);
          // weights: 3-D map (element, row of the N tile, chunk-tap index); one box = this CTA's rows of TS taps
          int wtap = (walk.nb * sg.nchunks) * taps;
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            mbar_wait(&a_empty[ia], pa ^ 1);
            if (leader) {
              // "full" barriers live in the leader CTA and count the bytes of both CTAs' loads
              if (lead_cta) mbar_arrive_expect_tx(&a_full[ia], a_bytes * G * S);
              uint32_t bar = 0;
              if constexpr (pair) bar = mapa_rank(smem_u32(&a_full[ia]), 0);
              for (int g = 0; g < G; ++g) {
                const ItemCoord cg = decode_tile(p, base + g);
                uint8_t* dst = sA + ia * p.a_buf_bytes + g * p.a_halo_bytes;
                if constexpr (pair)
                  tma_load_5d_pair(dst, &sg.tmap_act, bar, ch * CE, cg.x0 - pad, cg.y0 - pad, cg.b, sg.slot);
                else
                  tma_load_5d(dst, &sg.tmap_act, &a_full[ia], ch * CE, cg.x0 - pad, cg.y0 - pad, cg.b, sg.slot);
              }
            }
            if (++ia == NA) {
              ia = 0;
              pa ^= 1;
            }
            if (resident && !first) continue;   // weights already in shared memory
            for (int tap0 = 0; tap0 < taps; tap0 += TS) {   // TS divides taps (conv_halo_plan)
              if (!resident) mbar_wait(&w_empty[iw], pw ^ 1);
              if (leader) {
                // every bulk-async instruction costs its issuing thread ~440 cycles whatever its size (measured,
                // tools/micro/tma_rate.cu): a weight stage is ONE box, not one per tap
                if (lead_cta) mbar_arrive_expect_tx(&w_full[iw], static_cast<uint32_t>(stage_bytes * S));
                uint8_t* dst = sW + iw * stage_bytes;
                if constexpr (pair)
                  tma_load_3d_pair(dst, &sg.tmap_w, mapa_rank(smem_u32(&w_full[iw]), 0), 0, static_cast<int>(crank) * w_rows, wtap + tap0);
                else
                  tma_load_3d(dst, &sg.tmap_w, &w_full[iw], 0, 0, wtap + tap0);
              }
              if (++iw == NW) {
                iw = 0;
                pw ^= 1;
              }
            }
            wtap += taps;
          }
        }
        if constexpr (EPI == EPI_FWD) epi_fwd_loads<E>(p, sE, e_full, e_empty, walk, base, leader, es, eph);
        first = false;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (p.nseg > 0 && lead_cta) {
      const bool leader = elect_one();
      int ia = 0, iw = 0;
      uint32_t pa = 0, pw = 0;
      int abuf = 0;
      uint32_t aphase = 0;
      bool first = true;
      for (int base = walk.first_tile; base < walk.tiles_padded; base += walk.tile_stride) {
        mbar_wait(&tempty_bar[abuf], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(abuf * p.acc_cols);
        uint32_t accumulate = 0;
        if (resident) iw = 0;   // stage index = position inside the tile's K walk
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
              const int nt = TS;   // TS divides taps (conv_halo_plan)
              if (!resident || first) {
                mbar_wait(&w_full[iw], pw);
                tc_fence_after();
              }
              uint64_t bdesc = make_smem_desc_sw64(smem_u32(sW + iw * stage_bytes), 16, 512);
              for (int j = 0; j < nt; ++j, bdesc += static_cast<uint64_t>(w_bytes >> 4)) {
                // tap (dy, dx): the halo buffer seen through a row-shifted descriptor
                uint64_t adesc = adesc0 + static_cast<uint64_t>((dy * hw + dx) * (kChunkBytes >> 4));
                if (leader && !(p.debug_flags & 2)) {
                  for (int g = 0; g < G; ++g, adesc += static_cast<uint64_t>(p.a_halo_bytes >> 4)) {
                    if constexpr (pair) {
                      umma_pair<DT>(d_tmem + g * p.n_tile, adesc, bdesc, p.idesc, accumulate);
                      umma_pair<DT>(d_tmem + g * p.n_tile, adesc + 2, bdesc + 2, p.idesc, 1u);
                    } else {
                      umma<DT>(d_tmem + g * p.n_tile, adesc, bdesc, p.idesc, accumulate);
                      umma<DT>(d_tmem + g * p.n_tile, adesc + 2, bdesc + 2, p.idesc, 1u);
                    }
                  }
                }
                accumulate = 1;
                if (++dx == ks) {
                  dx = 0;
                  ++dy;
                }
              }
              if (leader && !resident) {
                if constexpr (pair) umma_commit_pair(&w_empty[iw]); else umma_commit(&w_empty[iw]);
              }
              if (++iw == NW) {
                iw = 0;
                pw ^= 1;
              }
            }
            if (leader) {
              if constexpr (pair) umma_commit_pair(&a_empty[ia]); else umma_commit(&a_empty[ia]);
            }
            if (++ia == NA) {
              ia = 0;
              pa ^= 1;
            }
          }
        }
        if (leader) {
          if constexpr (pair) umma_commit_pair(&tfull_bar[abuf]); else umma_commit(&tfull_bar[abuf]);
        }
        if (++abuf == p.n_acc) {
          abuf = 0;
          aphase ^= 1;
        }
        first = false;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ epilogue TMA stores
    // forward: warps 2 and 3 both store (alternate channel groups; 3 boxes per group), c_{t-1} is prefetched by
    // warp 0.  backward: warp 2 stores (2 boxes per group), warp 3 loads (4 boxes per group).
    if constexpr (EPI == EPI_FWD) epi_storer<E, EPI>(p, sE, st_ready, e_empty, walk, 0, 2);
    if constexpr (EPI == EPI_BWD) epi_storer<E, EPI>(p, sE, st_ready, e_empty, walk, 0, 1);
  } else if (warp == 3) {
    if constexpr (EPI == EPI_FWD) epi_storer<E, EPI>(p, sE, st_ready, e_empty, walk, 1, 2);
    if constexpr (EPI == EPI_BWD) epi_loader<E, EPI>(p, sE, e_full, e_empty, walk);
  } else {
    // ------------------------------------------------------------------ epilogue math (warps 4-11)
    // the epilogue of either CTA releases the accumulator buffer on the LEADER's "tmem empty" barrier
    uint32_t tempty_remote = 0;
    if constexpr (pair) tempty_remote = mapa_rank(smem_u32(&tempty_bar[0]), 0);
    if constexpr (EPI == EPI_RAW)
      epi_raw(p, warp, lane, tmem_base, tfull_bar, tempty_bar, walk, tempty_remote);
    else
      epi_math<E, EPI>(p, warp, lane, tmem_base, sE, e_full, st_ready, tfull_bar, tempty_bar, s_bias, s_headw, walk,
                       tempty_remote);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer still uses the pair's resources
  if (warp == 2) {
    tc_fence_after();
    if constexpr (pair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// shared-memory plan: epilogue I/O stages, tiles per group, taps per weight stage, stage / halo-buffer counts
int conv_halo_plan(int epi, int dtype, ConvGemmParams& p) {
  const int esize = dtype == NINT_BF16 ? 2 : 4;
  // ---- epilogue stage layout (nint_epilogue.cuh): [gates][c / c_t][c_{t-1}][dc] or [gates][c][h]
  const int gate_bytes = kTilePixels * 64 * esize;
  if (epi == EPI_FWD) {
    const int g = p.slot_g >= 0 ? gate_bytes : 0;
    p.e_off_c = g;
    p.e_off_h = g + kEpiBoxBytes16;
    p.e_off_c2 = p.e_off_dc = 0;
    p.e_stage_bytes = (p.e_off_h + kTilePixels * 16 * esize + 1023) & ~1023;
  } else if (epi == EPI_BWD) {
    p.e_off_c = gate_bytes;
    p.e_off_c2 = p.e_off_c + kEpiBoxBytes16;
    p.e_off_dc = p.e_off_c2 + kEpiBoxBytes16;
    p.e_off_h = 0;
    p.e_stage_bytes = p.e_off_dc + kEpiBoxBytes16;
  } else {
    p.e_off_c = p.e_off_c2 = p.e_off_dc = p.e_off_h = 0;
    p.e_stage_bytes = 0;
  }
  p.a_halo_bytes = halo_a_buf_bytes(p);
  if (p.a_halo_bytes == 0) p.a_halo_bytes = 1024;
  const int total = 227 * 1024 - 1024 - kHaloCtrlBytes - (4 * p.hc + p.hc) * 4 - 1024;
  const int w_bytes = (p.n_tile / p.cluster) * kChunkBytes;
  const int ns_max = epi == EPI_RAW ? 0 : 3, ns_min = epi == EPI_RAW ? 0 : 2;
  int max_k = 1;
  for (int s = 0; s < p.nseg; ++s) if (p.seg[s].ksize > max_k) max_k = p.seg[s].ksize;
  // ---- plan A: weights resident.  One stage per (segment, chunk) holding all taps; needs >= 3 halo buffers.
  if (p.nseg > 0) {
    int n_st = 0, ts = p.seg[0].ksize * p.seg[0].ksize;
    bool same_k = true;
    for (int s = 0; s < p.nseg; ++s) {
      n_st += p.seg[s].nchunks;
      same_k = same_k && (p.seg[s].ksize * p.seg[s].ksize == ts);
    }
    const long long res_bytes = static_cast<long long>(n_st) * ts * w_bytes;
    if (same_k && n_st <= kHaloMaxW) {
      for (int ns = ns_max; ns >= ns_min; --ns) {
        const long long rem = static_cast<long long>(total) - ns * p.e_stage_bytes - res_bytes;
        if (rem < 3LL * p.a_halo_bytes) continue;
        p.w_resident = 1;
        p.e_stages = ns;
        p.group = 1;
        p.a_buf_bytes = p.a_halo_bytes;
        p.na_bufs = static_cast<int>(rem / p.a_halo_bytes) > 6 ? 6 : static_cast<int>(rem / p.a_halo_bytes);
        p.taps_per_stage = ts;
        p.num_stages = n_st;
        p.acc_cols = p.n_tile;
        p.n_acc = 512 / p.acc_cols > kMaxAcc ? kMaxAcc : 512 / p.acc_cols;
        return 0;
      }
    }
  }
  // ---- plan B: weights stream through a ring of stages.
  p.w_resident = 0;
  // tiles per group: fill one 256-column accumulator buffer, at most 4 (2 for the memory-bound backward
  // epilogue, whose stages need the shared memory more than the weight stages need the reuse)
  int gmax = 256 / p.n_tile;
  if (gmax > (epi == EPI_BWD ? 2 : 4)) gmax = (epi == EPI_BWD ? 2 : 4);
  if (gmax < 1 || p.nseg == 0) gmax = 1;
  // three epilogue stages keep loads, math and stores of consecutive channel groups overlapped; fall back
  // to two when the operands would not get 2 halo buffers + 3 single-tap weight stages otherwise
  int NS = -1, G = 1;
  for (int ns = ns_max; ns >= ns_min && NS < 0; --ns)
    for (int g = gmax; g >= 1; --g)
      if (total - ns * p.e_stage_bytes >= 2 * g * p.a_halo_bytes + 3 * w_bytes) {
        NS = ns;
        G = g;
        break;
      }
  if (NS < 0) return 1;
  p.e_stages = NS;
  p.group = G;
  p.a_buf_bytes = G * p.a_halo_bytes;
  p.acc_cols = G * p.n_tile;
  p.n_acc = 512 / p.acc_cols > kMaxAcc ? kMaxAcc : 512 / p.acc_cols;
  const int budget = total - NS * p.e_stage_bytes;
  // every barrier round trip of the MMA-issuing thread costs a few hundred cycles: group taps so that
  // one weight stage carries plenty of tensor work, keep >= 3 weight stages in flight and spend the rest
  // of shared memory on halo buffers (a halo chunk is 180+ scattered 64-byte rows: ~1.5 us to arrive)
  int want = (1536 + p.n_tile - 1) / p.n_tile;
  if (want > max_k * max_k) want = max_k * max_k;
  int ts = 1;
  for (int cand = want; cand >= 1; --cand) {   // largest tap group that leaves room for 3 stages + 2 halo buffers and divides every tap count
    if (budget - 3 * cand * w_bytes < 2 * p.a_buf_bytes) continue;
    bool divides = true;
    for (int s = 0; s < p.nseg; ++s) divides = divides && ((p.seg[s].ksize * p.seg[s].ksize) % cand == 0);
    if (divides) { ts = cand; break; }
  }
  int nw = 3;
  int na = (budget - nw * ts * w_bytes) / p.a_buf_bytes;
  if (na > 6) na = 6;
  if (na < 1) na = 1;
  nw = (budget - na * p.a_buf_bytes) / (ts * w_bytes);
  if (nw > kHaloMaxW) nw = kHaloMaxW;
  if (nw < 1) return 1;
  p.na_bufs = na;
  p.taps_per_stage = ts;
  p.num_stages = nw;
  return 0;
}

template <typename E, int EPI, bool PAIR>
static cudaError_t launch_h(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  const int smem = conv_halo_smem_bytes(p);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<E, EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  constexpr int S = PAIR ? 2 : 1;
  const int tiles = p.B * p.tiles_x * p.tiles_y;
  const int groups = (tiles + S * p.group - 1) / (S * p.group);
  int cpn = num_sms / (S * p.n_blocks);   // clusters per n-block
  if (cpn > groups) cpn = groups;
  if (cpn <= 0) return cudaErrorInvalidConfiguration;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cpn * p.n_blocks * S);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PAIR ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv_halo_kernel<E, EPI, PAIR>, p);
}

template <typename E, bool PAIR>
static cudaError_t launch_e(int epi, const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (epi == EPI_FWD) return launch_h<E, EPI_FWD, PAIR>(p, num_sms, stream);
  if (epi == EPI_BWD) return launch_h<E, EPI_BWD, PAIR>(p, num_sms, stream);
  return launch_h<E, EPI_RAW, PAIR>(p, num_sms, stream);
}

cudaError_t launch_conv_halo(int epi, int dtype, const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.tile_w != 8 || p.tile_h != 16) return cudaErrorInvalidValue;
  if (p.cluster != 1 && p.cluster != 2) return cudaErrorInvalidValue;
  if (p.cluster == 2 && (p.n_tile % 32)) return cudaErrorInvalidValue;
  if (p.n_acc < 1 || p.n_acc > kMaxAcc || p.n_acc * p.acc_cols > kTmemCols) return cudaErrorInvalidValue;
  if (epi != EPI_RAW && (p.e_stages < 2 || p.e_stages > kEpiMaxStages)) return cudaErrorInvalidValue;
  if (dtype == NINT_BF16)
    return p.cluster == 2 ? launch_e<__nv_bfloat16, true>(epi, p, num_sms, stream)
                          : launch_e<__nv_bfloat16, false>(epi, p, num_sms, stream);
  return p.cluster == 2 ? launch_e<float, true>(epi, p, num_sms, stream) : launch_e<float, false>(epi, p, num_sms, stream);
}

}  // namespace nint
