// Epilogue shared by the implicit-GEMM convolution kernels: TMEM accumulator -> fused ConvLSTM
// pointwise math -> global memory.  Included by nint_conv_gemm.cu and nint_conv_halo.cu.
#pragma once
#include "nint_common.cuh"
#include "nint_kernels.h"

namespace nint {

constexpr int kConvThreads = 384;   // warps 0-3: producer / MMA / TMEM alloc / spare; warps 4-11: epilogue

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


// Runs on warps 4..11 of a CTA.  `tfull_bar` / `tempty_bar`: MMA <-> epilogue handshake of the two
// TMEM accumulator buffers (tempty expects one arrive per epilogue warp = 8).
template <typename E, int EPI>
__device__ __forceinline__ void conv_epilogue_loop(const ConvGemmParams& p, int warp, int lane, uint32_t tmem_base,
                                                   uint64_t* tfull_bar, uint64_t* tempty_bar, const float* s_bias,
                                                   const float* s_headw, int first_item, int item_stride,
                                                   int num_items_padded, int G = 1) {
  constexpr int DT = ElemTraits<E>::kDtype;
  constexpr bool FAST = (DT == NINT_BF16);
  const int num_items = p.n_blocks * p.B * p.tiles_x * p.tiles_y;
  {
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
    // one accumulator buffer (256 TMEM columns) holds the G consecutive tiles of an item group
    for (int base = first_item; base < num_items_padded; base += item_stride) {
      bool waited = (p.nseg == 0);
      auto wait_acc = [&]() {
        if (!waited) {
          mbar_wait(&tfull_bar[abuf], aphase);
          tc_fence_after();
          waited = true;
        }
      };
      for (int gi = 0; gi < G; ++gi) {
      const int item = base + gi;
      const ItemCoord c = decode_item(p, item);
      const int y = c.y0 + ty, x = c.x0 + tx;
      const bool valid = (item < num_items) && (ty < p.tile_h) && (y < p.H) && (x < p.W) && !(p.debug_flags & 1);
      const long long pix = (static_cast<long long>(c.b) * p.H + y) * p.W + x;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(abuf * 256 + gi * p.n_tile);
      if constexpr (EPI == EPI_FWD) {
        // model.py:221-229.  columns of this n-block: gate * hcb + cc
        const float* cprev = p.c_prev ? p.c_prev + pix * hc + c.nb * hcb : nullptr;
        float* cout = p.c_out + pix * hc + c.nb * hcb;
        E* hout = reinterpret_cast<E*>(p.h_out) + pix * p.hc_pad + c.nb * hcb;
        E* gout = p.gates_out ? reinterpret_cast<E*>(p.gates_out) + pix * 4 * hc + c.nb * p.n_tile : nullptr;
        const uint32_t bq = smem_u32(s_bias + c.nb * p.n_tile);
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
            for (int g = 0; g < 4; ++g) {   // bias: explicit 128-bit ld.shared (a generic pointer costs 64 LD.E)
              float bias[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) lds128(bq + (g * hcb + cg + j) * 4, &bias[j]);
#pragma unroll
              for (int j = 0; j < 16; ++j) a[g][j] += bias[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float gi = act_sigmoid<FAST>(a[0][j]);
              const float gf = act_sigmoid<FAST>(a[1][j]);
              const float gg = act_tanh<FAST>(a[2][j]);
              const float go = act_sigmoid<FAST>(a[3][j]);
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
          float hw[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) lds128(smem_u32(s_headw + c0 + j), &hw[j]);
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
              const float dhv = fmaf(dpred, hw[j], dh[j]);
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
      }  // tiles of the group
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
}

}  // namespace nint
