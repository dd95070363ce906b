// 3x3 convolution forward / data-gradient kernel for the 64 -> 64 channel layers of level 1 (conv12c, conv12e and
// their dgrads: the largest maps of the network): "three taps per instruction".
//
// With 64 output channels an M = 128, N = 64 tcgen05.mma costs ~54.5 cycles where its arithmetic needs 32, and it
// re-reads its 4 KB A tile from shared memory for 2 KB of B (igemm2.cuh runs these layers at ~70 % of that cap).
// Here the three taps (r, 0), (r, 1), (r, 2) of one kernel row share ONE instruction: B = [W(r,0); W(r,1); W(r,2)] is
// an N = 192 operand (three adjacent 8 KB weight slabs), A = the input pixels themselves, so the accumulator column
// block s at MMA pixel m holds x[m] * W(r, s), a contribution to the output pixel s columns to the LEFT of m:
//     out[y, q] = D0[q] + D1[q + 1] + D2[q + 2]        (D_s summed over r and K)
// 12 instructions of 96 cycles per tile and 64 input channels instead of 36 of 54.5, and 10 KB instead of 6 KB of
// operands per instruction (shared-memory traffic per output pixel drops by 35 %).
// The M = 128 tile is 4 rows x 32 columns of input pixels, so a warp of the epilogue owns one row and the
// shift-and-add across pixels is two warp shuffles per value; 30 of the 32 columns carry complete outputs.
//   A patch : 6 rows x 32 pixels x 64 channels (SWIZZLE_128B, dense: tap row r = byte offset r * 4096), one TMA box
//   B       : all 9 * K/64 weight slabs resident (loaded once per CTA, persistent over tiles)
//   D       : 2 accumulators x 192 fp32 columns in TMEM (epilogue of tile i overlaps the MMAs of tile i + 1)
//   out     : TMA store of a dense 4 x 30 pixel box (tails clipped by the tensor map)
//   fwd  : taps read x[y0 + r, x0 + s];  dgrad: dy[y0 + r - 2, x0 + s - 2] with rotated weights (pack mode 1), TMA zero
//   fill = the padding, optional fused ReLU-backward mask; optional fused 1x1 head (conv12e + finalconv).
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

struct alignas(64) Igemm3Params {
  CUtensorMap a_map;  // input view, box {64, 32, 6, 1}
  CUtensorMap b_map;  // packed weights [64][9*K], box {64, 64}
  CUtensorMap o_map;  // output view, box {64, 30, 4, 1}
  CUtensorMap m_map;  // ReLU-mask source view (geometry of the output), L2 prefetch only
  int kchunks;
  int tiles_x, tiles_y, n_img;
  int base_off;  // 0 or -2
  int n_a;
  const float* bias;
  int relu;
  const __nv_bfloat16* mask;
  long long mask_sn, mask_sy, mask_sx;
  int out_W, out_H;
  const float* head_w;  // fused 1x1 head, see igemm2.cuh
  const float* head_b;
  float* logits;
  uint8_t* argmax;
  int store_out;
  int dbg;   // experiment switches (ub_set_option 4): 1 = no pixel shift-and-add, 2 = no staging / store, 4 = TMEM read only
};

constexpr int IG3_MAX_A = 4, IG3_MAX_B = 18;
constexpr int IG3_A_BYTES = 6 * 32 * 128, IG3_B_BYTES = 64 * 128, IG3_E_BYTES = 16384;
constexpr int IG3_TW = 30, IG3_TH = 4;

template <bool HEAD>
__global__ void __launch_bounds__(384, 1) igemm3_kernel(const __grid_constant__ Igemm3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.n_a * IG3_A_BYTES;
  uint8_t* epi = sB + 9 * p.kchunks * IG3_B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(epi + 2 * IG3_E_BYTES);
  uint64_t* a_empty = a_full + IG3_MAX_A;
  uint64_t* b_full = a_empty + IG3_MAX_A;
  uint64_t* tfull = b_full + IG3_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(epi + 2 * IG3_E_BYTES + 1024);
  float* shead = reinterpret_cast<float*>(epi + 2 * IG3_E_BYTES + 2048);            // [2][64] + [2] (head only)
  float2* hpart = reinterpret_cast<float2*>(epi + 2 * IG3_E_BYTES + 2048 + 1024);   // [2 buffers][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_b = 9 * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    if (!HEAD || p.store_out) tma_prefetch_desc(&p.o_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.n_a; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < n_b; ++i) mbar_init(&b_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);   // 2 x 192 columns used
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int kchunks = p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform loop)
    int as = 0;
    uint32_t aph = 0;
    bool first = true;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int tx = t % p.tiles_x;
      int mt = t / p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * IG3_TW + p.base_off, y0 = ty * IG3_TH + p.base_off;
      if (p.mask != nullptr) {
        if (elect_one()) tma_prefetch_l2_4d(&p.m_map, 0, tx * IG3_TW, ty * IG3_TH, img);
        __syncwarp();
      }
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_empty[as], aph ^ 1, 30);
        if (elect_one()) {
          if ((p.dbg & 16) && !first) {
            mbar_arrive(&a_full[as]);                                      // experiment: no patch loads after the first tile
          } else {
            mbar_arrive_expect_tx(&a_full[as], (uint32_t)IG3_A_BYTES);
            tma_load_4d(sA + as * IG3_A_BYTES, &p.a_map, &a_full[as], kc * 64, x0, y0, img);
          }
        }
        __syncwarp();
        if (++as == p.n_a) {
          as = 0;
          aph ^= 1;
        }
        if (first) {
          if (elect_one()) {
            for (int tap = 0; tap < 9; ++tap) {
              const int slot = kc * 9 + tap;
              mbar_arrive_expect_tx(&b_full[slot], (uint32_t)IG3_B_BYTES);
              tma_load_2d(sB + slot * IG3_B_BYTES, &p.b_map, &b_full[slot], (tap * kchunks + kc) * 64, 0);
            }
          }
          __syncwarp();
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer (warp-uniform loop)
    constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 0, 0);
    const uint32_t hi = umma_desc_hi(1024);
    const uint32_t loA0 = umma_desc_lo(smem_u32(sA), 16), loB0 = umma_desc_lo(smem_u32(sB), 16);
    int as = 0;
    uint32_t aph = 0;
    int it = 0;
    bool resident = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1, 31);
      const uint32_t d0 = tmem_base + acc * 192;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_full[as], aph, 32);
        if (!resident)
          for (int tap = 0; tap < 9; ++tap) mbar_wait(&b_full[kc * 9 + tap], 0, 33);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t loA = loA0 + as * (IG3_A_BYTES >> 4);
          const uint32_t loB = loB0 + kc * 9 * (IG3_B_BYTES >> 4);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            if ((p.dbg & 8) && r > 0) break;                               // experiment: a third of the MMAs
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t aoff = (r * 32 * 128 + k * 32) >> 4;          // tap row r = 32 pixels further
              const uint32_t boff = (3 * r * IG3_B_BYTES + k * 32) >> 4;   // slabs (r,0), (r,1), (r,2) are adjacent
              umma_bf16(d0, umma_desc(hi, loA + aoff), umma_desc(hi, loB + boff), idesc, (kc > 0 || r > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[as]);
          if (kc == kchunks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++as == p.n_a) {
          as = 0;
          aph ^= 1;
        }
      }
      resident = true;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------ epilogue
    // warp w reads TMEM lanes 32*(w%4).. = tile row (w%4), lane = column; half h = (w-4)/4 of the 64 channels
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    const int srow = q * IG3_TW + lane;          // dense 4 x 30 staging box
    const bool col_ok = lane < IG3_TW;
    const bool mask_on = p.mask != nullptr;
    constexpr bool head_on = HEAD;
    const bool store_out = (!head_on || p.store_out != 0) && !(p.dbg & 2);
    if (et < 64 && p.bias != nullptr) sbias[et] = __ldg(p.bias + et);
    if (head_on) {
      if (et < 128) shead[et] = __ldg(p.head_w + et);
      if (et < 2) shead[128 + et] = __ldg(p.head_b + et);
    }
    named_bar_sync(1, 256);
    auto mask_row = [&](int tt) -> const __nv_bfloat16* {
      const int tx = tt % p.tiles_x;
      int mt = tt / p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int xx = tx * IG3_TW + lane, yy = ty * IG3_TH + q;
      if (!col_ok || xx >= p.out_W || yy >= p.out_H) return nullptr;
      return p.mask + (long long)img * p.mask_sn + (long long)yy * p.mask_sy + (long long)xx * p.mask_sx + h * 32;
    };
    uint4 mcur[4];
    bool mcur_valid = false;
    if (mask_on && (int)blockIdx.x < total_tiles) {
      const __nv_bfloat16* mp = mask_row(blockIdx.x);
      mcur_valid = mp != nullptr;
      if (mcur_valid) load_mask(mp, mcur);
    }
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int tx = t % p.tiles_x;
      int mt = t / p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * IG3_TW, y0 = ty * IG3_TH;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool px_valid = col_ok && (x0 + lane) < p.out_W && (y0 + q) < p.out_H;
      uint4 mnext[4];
      bool mnext_valid = false;
      if (mask_on && t + (int)gridDim.x < total_tiles) {   // next tile's mask row: one tile ahead of its use
        const __nv_bfloat16* mp = mask_row(t + gridDim.x);
        mnext_valid = mp != nullptr;
        if (mnext_valid) load_mask(mp, mnext);
      }
      mbar_wait(&tfull[acc], acc_phase, 34);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 192 + h * 32;
      uint32_t v0[32], v1[32], v2[32];
      tmem_ld_32x32b_x32(t_addr, v0);
      tmem_ld_32x32b_x32(t_addr + 64, v1);
      tmem_ld_32x32b_x32(t_addr + 128, v2);
      tmem_ld_wait();
      // all of this thread's TMEM reads of accumulator `acc` are complete: release it before the arithmetic
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (p.dbg & 4) continue;
      float f[32];
      if (p.dbg & 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v0[i]) + __uint_as_float(v1[i]) + __uint_as_float(v2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // out[q] = D0[q] + D1[q + 1] + D2[q + 2]: the neighbours' columns come over two shuffles
          const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[i]), 1);
          const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 2);
          f[i] = __uint_as_float(v0[i]) + a1 + a2;
        }
      }
      uint8_t* ebuf = epi + (it & 1) * IG3_E_BYTES;
      float2 hp = make_float2(0.0f, 0.0f);
      epi_finish(f, p.bias != nullptr ? sbias + h * 32 : nullptr, p.relu, mask_on, mcur_valid, mcur, ebuf, srow, h,
                 head_on ? shead + h * 32 : nullptr, &hp, store_out && col_ok);
      if (mask_on) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mcur[i] = mnext[i];
        mcur_valid = mnext_valid;
      }
      if (head_on && h == 1) hpart[(it & 1) * 128 + q * 32 + lane] = hp;
      fence_proxy_async_smem();
      if (et == 0 && store_out) tma_store_wait_read<0>();
      named_bar_sync(1, 256);
      if (et == 0 && store_out) {
        tma_store_4d(&p.o_map, ebuf, 0, x0, y0, img);
        tma_store_commit();
      }
      if (head_on && h == 0 && px_valid) {
        const float2 o = hpart[(it & 1) * 128 + q * 32 + lane];
        const float z0 = hp.x + o.x + shead[128], z1 = hp.y + o.y + shead[129];
        const long long pix = (long long)(y0 + q) * p.out_W + (x0 + lane);
        const long long plane = (long long)p.out_H * p.out_W;
        if (p.logits != nullptr) {
          p.logits[(long long)img * 2 * plane + pix] = z0;
          p.logits[((long long)img * 2 + 1) * plane + pix] = z1;
        }
        if (p.argmax != nullptr) p.argmax[(long long)img * plane + pix] = (z1 > z0) ? 1 : 0;
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ub
