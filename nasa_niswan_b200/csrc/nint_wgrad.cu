// Weight-gradient implicit GEMM on tcgen05 (sm_100a), batched over all T steps of BPTT.
//
//   dW[tap][q][col] += sum over (t, b, y, x)  dgates_t[b,y,x,q] * comb_t[b, y+dy-p, x+dx-p, col]
//
// which is what autograd's conv-weight backward computes for model.py:220 summed over the time
// loop (model.py:265); db[q] = sum dgates.  GEMM view: M = q (128 per m-block), N = ncols
// (x-part chunks then h-part chunks of the concatenated input), K = pixels.  Both operands are
// read "MN-major": a shared-memory panel is [128 pixel rows][32 channels] (64-byte rows with the
// 64-byte swizzle for bf16; 128-byte rows with the 128B/32B-atom swizzle for tf32, the only
// MN-major layout tcgen05 accepts for 32-bit operands), exactly what the channels-last TMA box
// delivers, so no transpose is ever materialised.
//
// One CTA owns (m-block, tap group, split): it keeps up to 512 fp32 accumulator columns in TMEM
// (taps_in_group x ncols [+ one chunk of columns for the bias]) across ALL its pixel tiles and
// flushes once at the end with fp32 atomics (split-K over pixel tiles and time).
// The bias gradient is an extra MMA against a constant panel of ones (group 0 only).
#include "nint_common.cuh"
#include "nint_kernels.h"
#include "nint_pair.cuh"

namespace nint {

constexpr int kWgThreads = 256;
constexpr int kWgCtrlBytes = 1024;
constexpr int kWgMaxBufs = 8;

constexpr int kWgPanelChannels = 32;  // channels per panel row, both dtypes
constexpr int kWgMPanels = 128 / kWgPanelChannels;
static inline int wg_panel_bytes(int dtype) { return kTilePixels * kWgPanelChannels * (dtype == NINT_BF16 ? 2 : 4); }

// halo mode: a B panel holds the (8+2p) x (16+2p) halo of the 8x16 pixel tile instead of 128 rows
int wgrad_b_panel_bytes(int dtype, int halo, int ksize) {
  if (!halo) return wg_panel_bytes(dtype);
  const int rows = (8 + (ksize & ~1)) * (16 + (ksize & ~1));
  return (rows * kWgPanelChannels * (dtype == NINT_BF16 ? 2 : 4) + 1023) & ~1023;
}
int wgrad_smem_bytes(int dtype, int bpanels, int a_bufs, int b_stages, int b_panel_bytes) {
  const int pb = wg_panel_bytes(dtype);
  return 1024 + a_bufs * kWgMPanels * pb + b_stages * bpanels * b_panel_bytes + pb + kWgCtrlBytes;
}
void wgrad_pick_buffers(int dtype, int bpanels, int b_panel_bytes, int* a_bufs, int* b_stages) {
  const int pb = wg_panel_bytes(dtype);
  const int budget = 227 * 1024 - 1024 - kWgCtrlBytes - pb;
  const int a1 = kWgMPanels * pb, b1 = bpanels * b_panel_bytes;
  int a = 2, b = (budget - a * a1) / b1;
  if (b < 2) {
    a = 1;
    b = (budget - a1) / b1;
  }
  if (b > kWgMaxBufs) b = kWgMaxBufs;
  *a_bufs = a;
  *b_stages = b;
}

template <typename E>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr int CE = kWgPanelChannels;          // channels per panel row
  constexpr int ROWB = CE * sizeof(E);          // 64 (bf16) / 128 (tf32) bytes per pixel row
  constexpr int PANEL = kTilePixels * ROWB;     // bytes per panel
  constexpr uint32_t LAYOUT = (DT == NINT_BF16) ? kLayoutSw64 : kLayoutSw128Base32;
  constexpr int MPANELS = kWgMPanels;           // A panels per m-block
  constexpr int ROWS_PER_MMA = 32 / sizeof(E);  // UMMA K: 16 (bf16) / 8 (tf32) pixel rows
  constexpr int KSTEPS = kTilePixels / ROWS_PER_MMA;
  constexpr uint32_t SBO = 512;                 // 8 rows x 64 B (bf16) / 4 rows x 128 B (tf32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bpanels = p.nchunks_b[0] + p.nchunks_b[1];
  const int a_buf_bytes = MPANELS * PANEL;
  const int b_stage_bytes = bpanels * p.b_panel_bytes;
  const int hpitch = 8 + (p.ksize & ~1);                        // halo mode: pixels per halo row
  const int hrows = hpitch * (16 + (p.ksize & ~1));
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.a_bufs * a_buf_bytes;
  uint8_t* sOnes = sB + p.b_stages * b_stage_bytes;
  uint8_t* ctrl = sOnes + PANEL;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + kWgMaxBufs;
  uint64_t* b_full = a_empty + kWgMaxBufs;
  uint64_t* b_empty = b_full + kWgMaxBufs;
  uint64_t* acc_full = b_empty + kWgMaxBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kWgCtrlBytes - 16);

  // blockIdx -> (split, group, m-block); CTAs of one split share their pixel tiles through L2
  int grp = 0;
  while (grp + 1 < p.n_groups && static_cast<int>(blockIdx.x) >= p.group_unit0[grp + 1]) ++grp;
  const int unit = static_cast<int>(blockIdx.x) - p.group_unit0[grp];
  const int mb = unit % p.m_blocks;
  const int split = unit / p.m_blocks;
  const int nsplits = p.group_splits[grp];
  const int tap_begin = p.group_tap0[grp];
  const int ntaps = p.group_tap0[grp + 1] - tap_begin;
  const bool do_bias = (grp == p.bias_group) && (p.db_acc != nullptr);
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.T * p.B * tiles_per_img;
  const int my_tiles = (total_tiles - split + nsplits - 1) / nsplits;  // split < splits <= total or 0 tiles
  const uint32_t panel_tx = static_cast<uint32_t>(p.tile_w * p.tile_h * ROWB);
  const int pad = p.ksize >> 1;

  // zero all operand panels once: rows >= tile_w*tile_h are never written by TMA and must not
  // contribute to the pixel reduction; fill the ones panel.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (p.a_bufs * a_buf_bytes + p.b_stages * b_stage_bytes) / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
    uint32_t one;
    if constexpr (DT == NINT_BF16) one = 0x3f803f80u; else one = 0x3f800000u;
    uint4* o = reinterpret_cast<uint4*>(sOnes);
    for (int i = threadIdx.x; i < PANEL / 16; i += kWgThreads) o[i] = make_uint4(one, one, one, one);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmap_dg);
    prefetch_tensormap(&p.tmap_b[0]);
    if (p.nchunks_b[1] > 0) prefetch_tensormap(&p.tmap_b[1]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgMaxBufs; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      const bool leader = elect_one();
      int ab = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int r = split + i * nsplits;
        const int tx = r % p.tiles_x;
        r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        r /= p.tiles_y;
        const int b = r % p.B;
        const int t = r / p.B;
        const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
        mbar_wait(&a_empty[ab], aph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&a_full[ab], panel_tx * MPANELS);
          for (int j = 0; j < MPANELS; ++j)
            tma_load_5d(sA + ab * a_buf_bytes + j * PANEL, &p.tmap_dg, &a_full[ab], mb * 128 + j * CE, x0, y0, b, t);
        }
        if (++ab == p.a_bufs) {
          ab = 0;
          aph ^= 1;
        }
        if (p.halo) {   // one halo load serves every tap of the group
          mbar_wait(&b_empty[bs], bph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&b_full[bs], static_cast<uint32_t>(hrows * ROWB * bpanels));
            uint8_t* dst = sB + bs * b_stage_bytes;
            // frame bank: the x tensor is [n_frames][H][W][C] and image b of step t is frame win_start[b] + t
            const int xb = p.win_start ? __ldg(p.win_start + b) + t : b, xs = p.win_start ? 0 : p.slot_b0[0] + t;
            for (int j = 0; j < p.nchunks_b[0]; ++j, dst += p.b_panel_bytes)
              tma_load_5d(dst, &p.tmap_b[0], &b_full[bs], p.chan0[0] + j * CE, x0 - pad, y0 - pad, xb, xs);
            for (int j = 0; j < p.nchunks_b[1]; ++j, dst += p.b_panel_bytes)
              tma_load_5d(dst, &p.tmap_b[1], &b_full[bs], p.chan0[1] + j * CE, x0 - pad, y0 - pad, b, p.slot_b0[1] + t);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        } else
        for (int ti = 0; ti < ntaps; ++ti) {
          const int tap = tap_begin + ti;
          const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
          mbar_wait(&b_empty[bs], bph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&b_full[bs], panel_tx * bpanels);
            uint8_t* dst = sB + bs * b_stage_bytes;
            const int xb = p.win_start ? __ldg(p.win_start + b) + t : b, xs = p.win_start ? 0 : p.slot_b0[0] + t;
            for (int j = 0; j < p.nchunks_b[0]; ++j, dst += PANEL)
              tma_load_5d(dst, &p.tmap_b[0], &b_full[bs], p.chan0[0] + j * CE, x0 + dx, y0 + dy, xb, xs);
            for (int j = 0; j < p.nchunks_b[1]; ++j, dst += PANEL)
              tma_load_5d(dst, &p.tmap_b[1], &b_full[bs], p.chan0[1] + j * CE, x0 + dx, y0 + dy, b, p.slot_b0[1] + t);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (my_tiles > 0) {
      const bool leader = elect_one();
      int ab = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&a_full[ab], aph);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_u32(sA + ab * a_buf_bytes), PANEL, SBO, LAYOUT);
        if (p.halo) {
          // k-step outer, tap inner: consecutive MMAs hit different accumulators (an accumulate chain
          // into one accumulator serialises on the MMA latency), every tap reads the halo in place
          // through a row-shifted MN-major descriptor.
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          // one K step = ROWS_PER_MMA pixels = 2 tile rows (bf16: 8-pixel groups one halo row apart) or
          // 1 tile row (tf32: 4-pixel groups 512 bytes apart)
          const uint32_t sbo_b = (DT == NINT_BF16) ? static_cast<uint32_t>(hpitch * ROWB) : 512u;
          const uint64_t bdesc0 = make_smem_desc(smem_u32(sB + bs * b_stage_bytes), p.b_panel_bytes, sbo_b, LAYOUT);
          if (leader) {
            const int reps = (NINT_DBG(p) & 2) ? 0 : ((NINT_DBG(p) & 4) ? 2 : 1);
            for (int rep = 0; rep < reps; ++rep)
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint64_t aoff = static_cast<uint64_t>((ks * ROWS_PER_MMA * ROWB) >> 4);
              const int row0 = ks * ROWS_PER_MMA / 8;
              const uint32_t acc = (i | ks) != 0 ? 1u : 0u;
              int dy = tap_begin / p.ksize, dx = tap_begin % p.ksize;
              for (int ti = 0; ti < ntaps; ++ti) {
                const uint64_t boff = static_cast<uint64_t>((((row0 + dy) * hpitch + dx) * ROWB) >> 4);
                umma<DT>(tmem_base + static_cast<uint32_t>(ti * p.acc_cols), adesc0 + aoff, bdesc0 + boff, p.idesc, acc);
                if (++dx == p.ksize) {
                  dx = 0;
                  ++dy;
                }
              }
              if (do_bias) {
                const uint64_t odesc0 = make_smem_desc(smem_u32(sOnes), PANEL, SBO, LAYOUT);
                umma<DT>(tmem_base + static_cast<uint32_t>(ntaps * p.acc_cols), adesc0 + aoff, odesc0 + aoff, p.idesc_bias, acc);
              }
            }
            umma_commit(&b_empty[bs]);
            umma_commit(&a_empty[ab]);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
          if (++ab == p.a_bufs) {
            ab = 0;
            aph ^= 1;
          }
          continue;
        }
        for (int ti = 0; ti < ntaps; ++ti) {
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint64_t bdesc0 = make_smem_desc(smem_u32(sB + bs * b_stage_bytes), PANEL, SBO, LAYOUT);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ti * p.acc_cols);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint64_t off = static_cast<uint64_t>((ks * ROWS_PER_MMA * ROWB) >> 4);
              umma<DT>(d_tmem, adesc0 + off, bdesc0 + off, p.idesc, (i | ks) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[bs]);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (do_bias && leader) {
          const uint64_t odesc0 = make_smem_desc(smem_u32(sOnes), PANEL, SBO, LAYOUT);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ntaps * p.acc_cols);
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint64_t off = static_cast<uint64_t>((ks * ROWS_PER_MMA * ROWB) >> 4);
            umma<DT>(d_tmem, adesc0 + off, odesc0 + off, p.idesc_bias, (i | ks) != 0 ? 1u : 0u);
          }
        }
        if (leader) umma_commit(&a_empty[ab]);
        if (++ab == p.a_bufs) {
          ab = 0;
          aph ^= 1;
        }
      }
      if (leader) umma_commit(acc_full);
    }
  } else if (warp >= 4 && my_tiles > 0) {
    // ------------------------------------------------------------------ epilogue: flush partial sums
    const int quad = warp & 3;
    const int q = mb * 128 + quad * 32 + lane;
    const bool valid = q < p.hc4;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // deterministic mode (dw_part): every split stores its partial sums into its own slice and a fixed-order
    // reduction follows (unpack_wgrad_kernel); default: fp32 atomics into dw_acc (summation order varies run to run)
    const long long part = static_cast<long long>(p.ksize) * p.ksize * p.hc4 * p.ncols * split;
    for (int ti = 0; ti < ntaps; ++ti) {
      const long long off = (static_cast<long long>(tap_begin + ti) * p.hc4 + q) * p.ncols + p.col0;
      float* dst = p.dw_part ? p.dw_part + part + off : p.dw_acc + off;
      for (int c0 = 0; c0 < p.acc_cols; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + ti * p.acc_cols + c0, v);
        tmem_ld_wait();
        if (valid) {
          if (p.dw_part) {
            store_elems<float, 16>(dst + c0, v);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + c0 + j, v[j]);
          }
        }
      }
    }
    if (do_bias) {
      float v[16];
      tmem_ld16(taddr + ntaps * p.acc_cols, v);
      tmem_ld_wait();
      if (valid) {
        if (p.db_part) p.db_part[static_cast<long long>(split) * p.hc4 + q] = v[0];
        else atomicAdd(p.db_acc + q, v[0]);
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

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant (bf16, 4*hc a multiple of 256): one tcgen05.mma.cta_group::2 covers M = 256 gate columns
// (128 per CTA) at the tensor pipe's nominal rate; the single-CTA M=128 MMA above needs ~100 cycles for the
// N=96 shape whose floor is 48 (tools/micro/mma_rate.cu).  Operands per pixel tile and CTA:
//   A  dgates  [128 px][128 q]  as 2 panels of 64 q   (128-byte rows, SWIZZLE_128B, MN-major)
//   B  comb    [halo px][N/2 channels] as panels of `b_pw` = 16 / 32 / 64 channels (32 / 64 / 128-byte rows with the
//      matching swizzle, MN-major): the pair MMA splits N between the CTAs, so the panel width is the largest that
//      divides N/2 and the x / h boundary (N/2 = 48 at hidden 64 -> 16; 128 + 128 channels -> 64).  Wide panels
//      matter: a TMA box row is one L2 request, and 16-channel rows are 32-byte requests.
// Warps: 0 A producer, 3 / 2 B producers (every bulk-async instruction occupies its issuing warp for ~450-750
// cycles; 2 also allocates TMEM), 1 / 8 MMA issuers (leader CTA, alternate tiles), 4-7 flush the accumulators with
// fp32 atomics.
constexpr int kWgPairMaxBStages = 4;
constexpr int kWgPairABufs = 3;   // the issuer awaits tile i+1 before issuing tile i: needs i-1, i, i+1 resident
constexpr int kWgPairAPanel = kTilePixels * 128;   // 64 q x 128 px bf16 = 16 KiB
constexpr int kWgPairThreads = 288;                // 9 warps: warp 8 is the second MMA issuer

__host__ __device__ static inline int wgp_b_panel_bytes(int ksize, int pw, int esize = 2) {
  const int rows = (8 + (ksize & ~1)) * (16 + (ksize & ~1));
  return (rows * pw * esize + 1023) & ~1023;
}
// bf16: three A buffers of 2 x 16 KiB panels; tf32: two A buffers of 4 x 16 KiB panels (32 q x 4 B rows)
static inline int wgp_smem_bytes(int half_cols, int ksize, int pw, int b_stages, int esize = 2) {
  const int a_bytes = esize == 2 ? kWgPairABufs * 2 * kWgPairAPanel : 2 * 4 * kWgPairAPanel;
  return 1024 + a_bytes + b_stages * (half_cols / pw) * wgp_b_panel_bytes(ksize, pw, esize) + 1024 + kWgCtrlBytes;
}
// widest B panel (channels) that tiles both CTAs' halves of N and does not straddle the x / h boundary
int wgrad_pair_panel_width(int cx_pad, int ncols) {
  for (int pw = 64; pw > 16; pw >>= 1)
    if ((ncols / 2) % pw == 0 && cx_pad % pw == 0) return pw;
  return 16;
}
// B stages that fit next to the three A buffers (0: the pair kernel cannot run this layer)
int wgrad_pair_b_stages(int cx_pad, int ncols, int ksize) {
  const int pw = wgrad_pair_panel_width(cx_pad, ncols);
  for (int st = kWgPairMaxBStages; st >= 2; --st)
    if (wgp_smem_bytes(ncols / 2, ksize, pw, st) <= 227 * 1024) return st;
  return 0;
}
// tf32 variant: 32-channel fp32 panels (SWIZZLE_128B_ATOM_32B), N rounded up to two whole panels per CTA
int wgrad_pair_b_stages_tf32(int n_mma, int ksize) {
  for (int st = kWgPairMaxBStages; st >= 2; --st)
    if (wgp_smem_bytes(n_mma / 2, ksize, 32, st, 4) <= 227 * 1024) return st;
  return 0;
}

template <int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgPairThreads, 1)
wgrad_pair_kernel(const __grid_constant__ WgradParams p) {
  // bf16: A = 2 panels of 64 q (SWIZZLE_128B), K step = 16 pixels = two tile rows.  tf32: A = 4 panels of 32 q and
  // B = 32-channel panels, both SWIZZLE_128B_ATOM_32B with 4-pixel groups 512 B apart (as the single-CTA kernel),
  // K step = 8 pixels = one tile row, twice as many MMAs at the same cycles each.
  constexpr bool BF = DT == NINT_BF16;
  constexpr int A_PANELS = BF ? 2 : 4;
  constexpr int A_PANEL_COLS = BF ? 64 : 32;
  constexpr int KSTEPS = BF ? 8 : 16;
  constexpr uint32_t A_KSTEP16 = BF ? 128u : 64u;      // A advance per K step in 16-byte units
  constexpr int ESIZE = BF ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool lead_cta = crank == 0;
  const int pw = p.b_pw;                                      // channels per B panel (16 / 32 / 64)
  const int rowb = pw * ESIZE;                                // bytes per pixel row of a B panel
  const int bp_all = p.acc_cols / pw;                         // panels of the concatenated input (tf32: + a zero panel)
  const int bp_cta = bp_all / 2;                              // this CTA's half of N
  const int bx16 = p.nchunks_b[0] * 32 / pw;                  // panels that come from the x tensor
  const int b_panel = wgp_b_panel_bytes(p.ksize, pw, ESIZE);
  const int nbs = p.b_stages;
  const int nab = p.a_bufs;
  constexpr int a_buf_bytes = A_PANELS * kWgPairAPanel;
  const int b_stage_bytes = bp_cta * b_panel;
  const int hpitch = 8 + (p.ksize & ~1);
  const int hrows = hpitch * (16 + (p.ksize & ~1));
  uint8_t* sA = smem;
  uint8_t* sB = sA + nab * a_buf_bytes;
  uint8_t* sOnes = sB + nbs * b_stage_bytes;
  uint8_t* ctrl = sOnes + 1024;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + kWgMaxBufs;
  uint64_t* b_full = a_empty + kWgMaxBufs;
  uint64_t* b_empty = b_full + kWgMaxBufs;
  uint64_t* acc_full = b_empty + kWgMaxBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kWgCtrlBytes - 16);
  volatile uint32_t* mma_issued = reinterpret_cast<volatile uint32_t*>(ctrl + kWgCtrlBytes - 32);

  // cluster -> (split, tap group, m-block of 256 q)
  const int cid = blockIdx.x >> 1;
  int grp = 0;
  while (grp + 1 < p.n_groups && cid >= p.group_unit0[grp + 1]) ++grp;
  const int unit = cid - p.group_unit0[grp];
  const int mb = unit % p.m_blocks;
  const int split = unit / p.m_blocks;
  const int nsplits = p.group_splits[grp];
  const int tap_begin = p.group_tap0[grp];
  const int ntaps = p.group_tap0[grp + 1] - tap_begin;
  const bool do_bias = (grp == p.bias_group) && (p.db_acc != nullptr);
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.T * p.B * tiles_per_img;
  const int my_tiles = (total_tiles - split + nsplits - 1) / nsplits;
  const int pad = p.ksize >> 1;

  {
    // the tiles are always full 8 x 16 boxes (TMA zero-fills outside the image), so no operand row is ever stale;
    // the ones panel feeds the bias-gradient MMA
    uint4* o = reinterpret_cast<uint4*>(sOnes);
    for (int i = threadIdx.x; i < 1024 / 16; i += kWgPairThreads) o[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmap_dg);
    prefetch_tensormap(&p.tmap_b[0]);
    prefetch_tensormap(&p.tmap_b[1]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgMaxBufs; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(acc_full, my_tiles >= 2 ? 2 : 1);   // one final commit per issuing warp that had a tile
    *mma_issued = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto tile_coords = [&](int i, int& x0, int& y0, int& b, int& t) {
    int r = split + i * nsplits;
    const int tx = r % p.tiles_x;
    r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    r /= p.tiles_y;
    b = r % p.B;
    t = r / p.B;
    x0 = tx * p.tile_w;
    y0 = ty * p.tile_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer: this CTA's 128 gate columns
    const bool leader = elect_one();
    int ab = 0;
    uint32_t aph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int x0, y0, b, t;
      tile_coords(i, x0, y0, b, t);
      mbar_wait(&a_empty[ab], aph ^ 1);
      if (leader) {
        if (lead_cta) mbar_arrive_expect_tx(&a_full[ab], static_cast<uint32_t>(2 * a_buf_bytes));
        const uint32_t bar = mapa_rank(smem_u32(&a_full[ab]), 0);
        const int q0 = mb * 256 + static_cast<int>(crank) * 128;
#pragma unroll
        for (int j = 0; j < A_PANELS; ++j)
          tma_load_5d_pair(sA + ab * a_buf_bytes + j * kWgPairAPanel, &p.tmap_dg, bar, q0 + j * A_PANEL_COLS, x0, y0, b, t);
      }
      if (++ab == nab) {
        ab = 0;
        aph ^= 1;
      }
    }
  } else if (warp == 3 || warp == 2) {
    // ------------------------------------------------------------------ B producers: this CTA's half of the channels
    // (warp 3 takes the even panels, warp 2 the odd ones)
    const bool leader = elect_one();
    const int jpar = warp == 3 ? 0 : 1;
    int bs = 0;
    uint32_t bph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int x0, y0, b, t;
      tile_coords(i, x0, y0, b, t);
      mbar_wait(&b_empty[bs], bph ^ 1);
      if (leader) {
        if (lead_cta && jpar == 0) mbar_arrive_expect_tx(&b_full[bs], static_cast<uint32_t>(2 * bp_cta * hrows * rowb));
        const uint32_t bar = mapa_rank(smem_u32(&b_full[bs]), 0);
        uint8_t* dst = sB + bs * b_stage_bytes + jpar * b_panel;
        const int xb = p.win_start ? __ldg(p.win_start + b) + t : b, xs = p.win_start ? 0 : p.slot_b0[0] + t;   // frame bank
        for (int j = jpar; j < bp_cta; j += 2, dst += 2 * b_panel) {
          const int pj = static_cast<int>(crank) * bp_cta + j;   // 16-channel panel of the concatenated input
          if (pj < bx16)
            tma_load_5d_pair(dst, &p.tmap_b[0], bar, p.chan0[0] + pj * pw, x0 - pad, y0 - pad, xb, xs);
          else
            tma_load_5d_pair(dst, &p.tmap_b[1], bar, p.chan0[1] + (pj - bx16) * pw, x0 - pad, y0 - pad, b, p.slot_b0[1] + t);
        }
      }
      if (++bs == nbs) {
        bs = 0;
        bph ^= 1;
      }
    }
  } else if (warp == 1 || warp == 8) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA): two warps take
    // alternate pixel tiles.  The issuing thread only runs a few MMAs ahead of the tensor pipe, so with one issuer
    // the ~500 cycles of barrier work per tile were tensor idle time; with two, one warp's waits and commits hide
    // behind the other's MMA stream.  Tiles enter the pipe strictly in order (`mma_issued`): deterministic sums.
    if (my_tiles > 0 && lead_cta) {
      const int which = warp == 1 ? 0 : 1;
      const bool leader = elect_one();
      const uint32_t idesc = p.idesc, idesc_bias = p.idesc_bias;
      const uint32_t ncols = static_cast<uint32_t>(p.acc_cols);
      const int ksize = p.ksize;
      const bool issue_any = !(NINT_DBG(p) & 2);
      // (A-collector reuse across the taps of a K step -- tcgen05.mma .collector::a::fill / use / lastuse -- was tried and
      // is slower: the MMAs of a K step serialise on the collector.  Even as a run-time knob it cost this loop ~10 %,
      // so the code is gone: DESIGN.md section 6.)
      // MN-major descriptors.  A: 64-q atoms (128-byte rows, SWIZZLE_128B) one panel apart, 8-pixel groups 1 KiB apart.
      // B: pw-channel atoms (32 / 64 / 128-byte rows, matching swizzle) one panel apart, 8-pixel groups one halo row apart.
      const uint64_t adesc_base = make_smem_desc(0, kWgPairAPanel, BF ? 1024u : 512u, BF ? 2u : kLayoutSw128Base32);
      const uint64_t bdesc_base = make_smem_desc(0, static_cast<uint32_t>(b_panel),
                                                 BF ? static_cast<uint32_t>(hpitch * rowb) : 512u,
                                                 BF ? (pw == 16 ? 6u : (pw == 32 ? 4u : 2u)) : kLayoutSw128Base32);
      const uint64_t odesc = make_smem_desc(smem_u32(sOnes), 512, 256, 6);
      const uint32_t ahi = static_cast<uint32_t>(adesc_base >> 32), bhi = static_cast<uint32_t>(bdesc_base >> 32);
      const uint32_t alo_base = static_cast<uint32_t>(adesc_base), blo_base = static_cast<uint32_t>(bdesc_base);
      const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
      const int dy0 = tap_begin / ksize, dx0 = tap_begin % ksize;
      const uint32_t krow16 = static_cast<uint32_t>(((BF ? 2 : 1) * hpitch * rowb) >> 4);   // one K step = two (bf16) / one (tf32) halo rows
      uint32_t tap_off[5];
#pragma unroll
      for (int ti = 0; ti < 5; ++ti) {
        const int tp = tap_begin + (ti < ntaps ? ti : 0);
        tap_off[ti] = static_cast<uint32_t>((((tp / ksize) * hpitch + tp % ksize) * rowb) >> 4);
      }
      int last_mine = -1;
      for (int i = which; i < my_tiles; i += 2) {
        const int ab = i % nab, bs = i % nbs;
        const uint32_t aph = static_cast<uint32_t>(i / nab) & 1u, bph = static_cast<uint32_t>(i / nbs) & 1u;
        mbar_wait(&a_full[ab], aph);
        mbar_wait(&b_full[bs], bph);
        spin_until_at_least(mma_issued, static_cast<uint32_t>(i));   // the other warp has issued all of tile i-1
        tc_fence_after();
        if (leader && issue_any) {
          const uint32_t a0 = alo_base + sA16 + static_cast<uint32_t>((ab * a_buf_bytes) >> 4);
          const uint32_t b0 = blo_base + sB16 + static_cast<uint32_t>((bs * b_stage_bytes) >> 4);
          uint32_t acc = i != 0 ? 1u : 0u;
          const int reps = (NINT_DBG(p) & 4) ? 2 : 1;   // experiment: issue every tile's MMA stream twice
          for (int rep = 0; rep < reps; ++rep)
          if (ntaps <= 5) {
            // per-tap B offsets are loop invariants: the MMA stream is adds + tcgen05.mma only
#pragma unroll 2
            for (int ks = 0; ks < KSTEPS; ++ks) {
              // K step = 16 pixels = 2 tile rows: A advances 16 x 128 B, B two halo rows (tf32: 8 pixels, one row)
              const uint32_t alo = a0 + static_cast<uint32_t>(ks) * A_KSTEP16;
              const uint32_t bk = b0 + static_cast<uint32_t>(ks) * krow16;
              uint32_t d = tmem_base;
#pragma unroll
              for (int ti = 0; ti < 5; ++ti) {
                if (ti < ntaps) umma_lohi<DT, true>(d, alo, ahi, bk + tap_off[ti], bhi, idesc, acc);
                d += ncols;
              }
              if (BF && do_bias)
                umma_lohi<DT, true>(tmem_base + static_cast<uint32_t>(ntaps) * ncols, alo, ahi,
                                           static_cast<uint32_t>(odesc), static_cast<uint32_t>(odesc >> 32), idesc_bias, acc);
              acc = 1;
            }
          } else {
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint32_t alo = a0 + static_cast<uint32_t>(ks) * A_KSTEP16;
              uint32_t blo = b0 + static_cast<uint32_t>(((((BF ? 2 : 1) * ks + dy0) * hpitch + dx0) * rowb) >> 4);
              uint32_t d = tmem_base;
              int dx = dx0;
              for (int ti = 0; ti < ntaps; ++ti) {
                umma_lohi<DT, true>(d, alo, ahi, blo, bhi, idesc, acc);
                d += ncols;
                blo += static_cast<uint32_t>(rowb >> 4);
                if (++dx == ksize) {
                  dx = 0;
                  blo += static_cast<uint32_t>(((hpitch - ksize) * rowb) >> 4);
                }
              }
              if (BF && do_bias)
                umma_lohi<DT, true>(d, alo, ahi, static_cast<uint32_t>(odesc), static_cast<uint32_t>(odesc >> 32),
                                           idesc_bias, acc);
              acc = 1;
            }
          }
        }
        if (leader) {
          __threadfence_block();
          *mma_issued = static_cast<uint32_t>(i + 1);
        }
        __syncwarp();
        umma_commit_elect<true>(&b_empty[bs]);
        umma_commit_elect<true>(&a_empty[ab]);
        last_mine = i;
      }
      if (last_mine >= 0) umma_commit_elect<true>(acc_full);
    }
  } else if (warp >= 4 && warp < 8 && my_tiles > 0) {
    // ------------------------------------------------------------------ epilogue: flush this CTA's 128 gate columns
    const int quad = warp & 3;
    const int q = mb * 256 + static_cast<int>(crank) * 128 + quad * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const long long part = static_cast<long long>(p.ksize) * p.ksize * p.hc4 * p.ncols * split;   // deterministic mode
    for (int ti = 0; ti < ntaps; ++ti) {
      const long long off = (static_cast<long long>(tap_begin + ti) * p.hc4 + q) * p.ncols + p.col0;
      float* dst = p.dw_part ? p.dw_part + part + off : p.dw_acc + off;
      for (int c0 = 0; c0 < p.real_cols; c0 += 16) {   // columns beyond real_cols are the tf32 variant's zero panel
        float v[16];
        tmem_ld16(taddr + ti * p.acc_cols + c0, v);
        tmem_ld_wait();
        if (p.dw_part) {
          store_elems<float, 16>(dst + c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + c0 + j, v[j]);
        }
      }
    }
    if (do_bias) {
      float v[16];
      tmem_ld16(taddr + ntaps * p.acc_cols, v);
      tmem_ld_wait();
      if (p.db_part) p.db_part[static_cast<long long>(split) * p.hc4 + q] = v[0];
      else atomicAdd(p.db_acc + q, v[0]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

int wgrad_pair_supported(int dtype, int hc4, int cx_pad, int ncols, int ksize) {
  if (dtype != NINT_BF16 || (hc4 % 256) != 0 || (ncols % 32) != 0 || ncols > 256) return 0;
  return wgrad_pair_b_stages(cx_pad, ncols, ksize) >= 2;
}

template <int DT>
static cudaError_t launch_wg_pair(const WgradParams& p, cudaStream_t stream) {
  const int smem = wgp_smem_bytes(p.acc_cols / 2, p.ksize, p.b_pw, p.b_stages, DT == NINT_BF16 ? 2 : 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_pair_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = 2 * p.group_unit0[p.n_groups];
  if (grid <= 0) return cudaSuccess;
  wgrad_pair_kernel<DT><<<grid, kWgPairThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template <typename E>
static cudaError_t launch_wg(const WgradParams& p, cudaStream_t stream) {
  const int dtype = ElemTraits<E>::kDtype;
  const int smem = wgrad_smem_bytes(dtype, p.nchunks_b[0] + p.nchunks_b[1], p.a_bufs, p.b_stages, p.b_panel_bytes);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = p.group_unit0[p.n_groups];
  if (grid <= 0) return cudaSuccess;
  wgrad_kernel<E><<<grid, kWgThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_wgrad(int dtype, const WgradParams& p, cudaStream_t stream) {
  if (p.pair) return dtype == NINT_BF16 ? launch_wg_pair<NINT_BF16>(p, stream) : launch_wg_pair<NINT_TF32>(p, stream);
  if (dtype == NINT_BF16) return launch_wg<__nv_bfloat16>(p, stream);
  return launch_wg<float>(p, stream);
}

}  // namespace nint
