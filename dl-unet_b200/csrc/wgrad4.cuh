// 3x3 weight-gradient kernel for layers with 64 OUTPUT channels (conv12c, conv11e, conv12e).
//
//   dW[co, (r,s), ci] = sum_{pixels (y,x)} dY[y, x, co] * X[y + r, x + s, ci]
//
// With Cout = 64 an M = 128 accumulator is half empty in wgrad3/wgrad5.  Here the second half of M is
// filled with a ONE-PIXEL-SHIFTED copy of dY, so one 128 x 128 x 16 instruction (the cheapest shape that
// runs at the nominal tensor rate) produces three taps x 64 ci x 64 co:
//   A (MN-major, SWIZZLE_128B): block 0 = dY[y, x-1, :], block 1 = dY[y, x, :]     (LBO = one 128 B row)
//   B (MN-major, SWIZZLE_128B): block 0 = X[y+r, x, :],  block 1 = X[y+r, x+2, :]  (LBO = two rows)
//   D[(blk, co), (sb, ci)] = sum_x A_blk[x, co] * B_sb[x, ci]
//       blk 1, sb 0 -> tap s = 0      blk 1, sb 1 -> tap s = 2
//       blk 0, sb 0 -> tap s = 1      blk 0, sb 1 -> tap s = 3 (discarded)
// i.e. 75 % useful work at 64 cycles, against 50 % at 56 cycles per 32-channel tile before.
// The dY box is 17 pixels wide and starts at x0-1 (out-of-bounds = zero), the X patch is 18 x (KH+2);
// chunks cover x in [0, W] so that the shifted copy also sees the last column.  One CTA owns a
// 64-channel ci tile and all nine taps (three accumulators of 128 columns), split-K over pixel chunks.
//
// p.pack (round 2, default): nine taps in TEN slots instead of twelve.  The dY box gets one more row on top (17 x (KH+1)
// pixels from (x0-1, y0-1)), so the second half of M can also be a ONE-ROW-shifted copy, and per 16-pixel K step
//   MMA 1 (N = 192): A = [dY(y, x-1), dY(y, x)] (LBO one pixel),    B = [X(y, x), X(y+1, x), X(y+2, x)] (LBO one patch row)
//                    -> taps (rb, 1) from block 0 and (rb, 0) from block 1, rb = 0..2: six taps, no waste
//   MMA 2 (N = 128): A = [dY(y-1, x), dY(y, x)] (LBO one box row), B = [X(y, x+2), X(y+2, x+2)]    (LBO two patch rows)
//                    -> taps (0, 2), (2, 2) from block 1, (1, 2) and a discarded (3, 2) from block 0
// 96 + 64 = 160 tensor cycles per K step against 3 x 64 = 192; chunks also cover y in [0, H] so that the row-shifted copy
// sees the last row.
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) Wgrad4Params {
  CUtensorMap g_map;  // dY view, box {64, 17, KH, 1}
  CUtensorMap x_map;  // X  view, box {64, 18, KH+2, 1}
  int KH;
  int chunks_x, chunks_y, n_img;
  int ci_tiles, splits;
  int Cin;
  int nstage, stage_bytes, g_bytes;
  int per;                // pixel chunks per split (no split is empty)
  int pack;               // 1: ten-slot tap packing (dY box has KH+1 rows and starts at y0-1)
  float* dw;              // slabs [splits][Cout][9][Cin] fp32: split s STORES its partial sums into slab s (deterministic;
                          // the slabs are summed in fixed order by the unpack kernel)
  long long slab_stride;  // elements between slabs
};

struct Wgrad4Cfg {
  static constexpr int MAX_STAGES = 8;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 232448 - 1024 - BAR_BYTES;
  static constexpr int TMEM_COLS = 512;  // 3 x 128 used
  static int g_bytes(int kh) { return ((17 * kh * 128 + 1023) / 1024) * 1024; }
  static int x_bytes(int kh) { return ((18 * (kh + 2) * 128 + 1023) / 1024) * 1024; }
  static int stage_bytes(int kh) { return g_bytes(kh) + x_bytes(kh); }
  static int nstage(int kh) {
    const int n = SMEM_LIMIT / stage_bytes(kh);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
  static int smem_bytes(int kh) { return nstage(kh) * stage_bytes(kh) + BAR_BYTES + 1024; }
};

__global__ void __launch_bounds__(256, 1) wgrad4_kernel(const __grid_constant__ Wgrad4Params p) {
  using C = Wgrad4Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NSTAGE = p.nstage, STAGE_BYTES = p.stage_bytes, G_BYTES = p.g_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int ct = blockIdx.x % p.ci_tiles;
  const int split = blockIdx.x / p.ci_tiles;
  const int total_chunks = p.n_img * p.chunks_y * p.chunks_x;
  const int per = p.per;
  const int c_begin = split * per;
  const int c_end = min(total_chunks, c_begin + per);
  const int n_iters = c_end - c_begin;
  const int KH = p.KH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.g_map);
    tma_prefetch_desc(&p.x_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_iters > 0) {
    if (warp == 0) {
      const uint32_t tx_bytes = 17u * (KH + (p.pack ? 1 : 0)) * 128u + 18u * (KH + 2) * 128u;
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int cx = c % p.chunks_x;
        const int r2 = c / p.chunks_x;
        const int cy = r2 % p.chunks_y;
        const int img = r2 / p.chunks_y;
        const int x0 = cx * 16, y0 = cy * KH;
        mbar_wait(&empty[stage], phase ^ 1, 40);
        if (elect_one()) {
          uint8_t* sg = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          tma_load_4d(sg, &p.g_map, &full[stage], 0, x0 - 1, p.pack ? y0 - 1 : y0, img);
          tma_load_4d(sg + G_BYTES, &p.x_map, &full[stage], ct * 64, x0, y0, img);
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);
      const uint32_t hiD = umma_desc_hi(1024);
      const uint32_t loA0 = umma_desc_lo(smem_u32(smem), 128);           // M blocks one pixel row apart
      const uint32_t loB0 = umma_desc_lo(smem_u32(smem) + G_BYTES, 256); // N blocks two pixel rows apart
      const uint32_t stage16 = (uint32_t)STAGE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_iters; ++i) {
        mbar_wait(&full[stage], phase, 41);
        tc_fence_after();
        if (p.pack) {
          if (elect_one()) {
            constexpr uint32_t idesc192 = umma_idesc_bf16(128, 192, 1, 1);
            const uint32_t so = stage * stage16;
            // box row 0 is dY row y0-1: the chunk's own rows start one box row (17 pixels) further down
            uint32_t loAx = umma_desc_lo(smem_u32(smem) + 17 * 128, 128) + so;          // [dY(y, x-1), dY(y, x)]
            uint32_t loAy = umma_desc_lo(smem_u32(smem) + 128, 17 * 128) + so;          // [dY(y-1, x), dY(y, x)]
            uint32_t loB3 = umma_desc_lo(smem_u32(smem) + G_BYTES, 18 * 128) + so;      // X(y + rb, x), rb = 0..2
            uint32_t loB2 = umma_desc_lo(smem_u32(smem) + G_BYTES + 256, 36 * 128) + so;  // X(y, x+2), X(y+2, x+2)
            for (int y = 0; y < KH; ++y) {
              const uint32_t acc = (i > 0 || y > 0) ? 1u : 0u;
              umma_bf16(tmem_base, umma_desc(hiD, loAx), umma_desc(hiD, loB3), idesc192, acc);
              umma_bf16(tmem_base + 192, umma_desc(hiD, loAy), umma_desc(hiD, loB2), idesc, acc);
              loAx += 136;
              loAy += 136;
              loB3 += 144;
              loB2 += 144;
            }
            umma_commit(&empty[stage]);
            if (i == n_iters - 1) umma_commit(tfull);
          }
        } else if (elect_one()) {
          uint32_t loA = loA0 + stage * stage16, loB = loB0 + stage * stage16;
          for (int y = 0; y < KH; ++y) {
            const uint32_t acc = (i > 0 || y > 0) ? 1u : 0u;
            umma_bf16(tmem_base, umma_desc(hiD, loA), umma_desc(hiD, loB), idesc, acc);              // r = 0
            umma_bf16(tmem_base + 128, umma_desc(hiD, loA), umma_desc(hiD, loB + 144), idesc, acc);  // r = 1
            umma_bf16(tmem_base + 256, umma_desc(hiD, loA), umma_desc(hiD, loB + 288), idesc, acc);  // r = 2
            loA += 136;  // 17 rows x 128 B >> 4
            loB += 144;  // 18 rows x 128 B >> 4
          }
          umma_commit(&empty[stage]);
          if (i == n_iters - 1) umma_commit(tfull);
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      const int m = q * 32 + lane;  // accumulator row: block m / 64, channel co = m % 64
      const int blk = m >> 6, co = m & 63;
      mbar_wait(tfull, 0, 42);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      if (p.pack) {
        // accumulator 1 (columns 0..191): block rb of 64 ci -> tap (rb, blk ? 0 : 1); accumulator 2 (192..319): block b ->
        // tap (2b, 2) for blk 1, (2b + 1, 2) for blk 0 (b = 1: discarded)
#pragma unroll 1
        for (int cb = 0; cb < 5; ++cb) {
          const int r = cb < 3 ? cb : (blk ? 2 * (cb - 3) : 2 * (cb - 3) + 1);
          const int s = cb < 3 ? (blk ? 0 : 1) : 2;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_addr + cb * 64 + hh * 32, v);
            tmem_ld_wait();
            if (r < 3) {
              float* drow = p.dw + (long long)split * p.slab_stride + ((long long)co * 9 + r * 3 + s) * p.Cin + ct * 64 + hh * 32;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 a = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                       __uint_as_float(v[i + 3]));
                __stcg(reinterpret_cast<float4*>(drow + i), a);
              }
            }
          }
        }
      }
#pragma unroll 1
      for (int r = 0; r < (p.pack ? 0 : 3); ++r) {
#pragma unroll 1
        for (int sb = 0; sb < 2; ++sb) {
          const int s = blk ? 2 * sb : 1 + 2 * sb;  // blk 1: taps 0, 2;  blk 0: taps 1, (3)
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_addr + r * 128 + sb * 64 + hh * 32, v);
            tmem_ld_wait();
            if (s < 3) {
              float* drow = p.dw + (long long)split * p.slab_stride + ((long long)co * 9 + r * 3 + s) * p.Cin + ct * 64 + hh * 32;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 a = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                       __uint_as_float(v[i + 3]));
                __stcg(reinterpret_cast<float4*>(drow + i), a);
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace ub
