// 3x3 weight-gradient kernel for the layers with >= 128 input channels ("one tap row per CTA, three
// s taps per instruction, N = 192").
//
//   dW[co, (r,s), ci] = sum_{pixels (y,x)} dY[y, x, co] * X[y + r, x + s, ci]
//
// wgrad_kernel (one tap per CTA, N = 256) runs its MMAs at full rate but moves 48 KB through L2 per
// 512 tensor cycles (94 B/clk/SM) and is L2-bound at ~45 % of the tensor peak; wgrad3_kernel (all nine
// taps, 32-channel tile) has little traffic but N = 96 instructions (86 % ceiling) and re-reads dY
// Cin/32 times.  This kernel sits between the two: a CTA owns (tap row r, 128-wide co tile, 128-wide
// ci group = two 64-channel blocks, K split).  Per 16 x KH pixel chunk it loads dY (box {64,16,KH} per
// co block) and, per ci block, the 18 x KH pixel rows y+r of X (box {64,18,KH}, SWIZZLE_128B).  Both
// operands are MN-major; the three s taps are the three N blocks of ONE instruction (leading-dimension
// offset = one 128 B pixel row), so N = 3 x 64 = 192 and the MMA runs at its nominal 96 cycles.
// Two accumulators (one per ci block) of 192 fp32 columns.  L2->SM: ~69 KB per 1536 tensor cycles
// (45 B/clk/SM); dY is re-read 3 * Cin/128 times, by CTAs adjacent in the grid.
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) Wgrad5Params {
  CUtensorMap g_map;  // dY view, box {64, 16, KH, 1}
  CUtensorMap x_map;  // X  view, box {64, 18, KH, 1}
  int KH;
  int chunks_x, chunks_y, n_img;
  int m_tiles, ci_groups, splits;
  int Cout, Cin;
  int nstage, stage_bytes, g_bytes, x_block_bytes;
  int per;                // pixel chunks per split (no split is empty)
  float* dw;              // slabs [splits][Cout][9][Cin] fp32: split s STORES its partial sums into slab s (deterministic;
                          // the slabs are summed in fixed order by the unpack kernel)
  long long slab_stride;  // elements between slabs
};

struct Wgrad5Cfg {
  static constexpr int MAX_STAGES = 8;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 232448 - 1024 - BAR_BYTES;
  static constexpr int TMEM_COLS = 512;  // 2 x 192 used
  static int g_bytes(int kh) { return 2 * 16 * kh * 128; }
  static int x_block_bytes(int kh) { return ((18 * kh * 128 + 1023) / 1024) * 1024; }
  static int stage_bytes(int kh) { return g_bytes(kh) + 2 * x_block_bytes(kh); }
  static int nstage(int kh) {
    const int n = SMEM_LIMIT / stage_bytes(kh);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
  static int smem_bytes(int kh) { return nstage(kh) * stage_bytes(kh) + BAR_BYTES + 1024; }
};

__global__ void __launch_bounds__(256, 1) wgrad5_kernel(const __grid_constant__ Wgrad5Params p) {
  using C = Wgrad5Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NSTAGE = p.nstage, STAGE_BYTES = p.stage_bytes, G_BYTES = p.g_bytes, XB = p.x_block_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // grid order: r fastest, then ci group, then co tile: CTAs that share a dY chunk run side by side
  int w = blockIdx.x;
  const int r = w % 3;
  w /= 3;
  const int cg = w % p.ci_groups;
  w /= p.ci_groups;
  const int mt = w % p.m_tiles;
  const int split = w / p.m_tiles;

  const int total_chunks = p.n_img * p.chunks_y * p.chunks_x;
  const int per = p.per;
  const int c_begin = split * per;
  const int c_end = min(total_chunks, c_begin + per);
  const int n_iters = c_end - c_begin;
  const int m_blocks = min(2, (p.Cout - mt * 128 + 63) / 64);
  const int n_blocks = min(2, (p.Cin - cg * 128) / 64);
  const int KH = p.KH;
  const uint32_t g_block_bytes = 16u * KH * 128u;

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
      const uint32_t tx_bytes = (uint32_t)m_blocks * g_block_bytes + (uint32_t)n_blocks * 18u * KH * 128u;
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int cx = c % p.chunks_x;
        const int r2 = c / p.chunks_x;
        const int cy = r2 % p.chunks_y;
        const int img = r2 / p.chunks_y;
        const int x0 = cx * 16, y0 = cy * KH;
        mbar_wait(&empty[stage], phase ^ 1, 30);
        if (elect_one()) {
          uint8_t* sg = smem + stage * STAGE_BYTES;
          uint8_t* sx = sg + G_BYTES;
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          for (int b = 0; b < m_blocks; ++b)
            tma_load_4d(sg + b * g_block_bytes, &p.g_map, &full[stage], mt * 128 + b * 64, x0, y0, img);
          for (int b = 0; b < n_blocks; ++b)
            tma_load_4d(sx + b * XB, &p.x_map, &full[stage], cg * 128 + b * 64, x0, y0 + r, img);
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 1, 1);
      // A: 16 pixel rows (two 8-row groups 1 KB apart) x 128 co (two 64-wide blocks g_block_bytes apart)
      // B: 16 pixel rows x 192 = three 64-wide blocks, block s = the patch shifted by s pixels (128 B)
      const uint32_t hiD = umma_desc_hi(1024);
      const uint32_t loA0 = umma_desc_lo(smem_u32(smem), g_block_bytes);
      const uint32_t loB0 = umma_desc_lo(smem_u32(smem) + G_BYTES, 128);
      const uint32_t stage16 = (uint32_t)STAGE_BYTES >> 4, xb16 = (uint32_t)XB >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_iters; ++i) {
        mbar_wait(&full[stage], phase, 31);
        tc_fence_after();
        if (elect_one()) {
          uint32_t loA = loA0 + stage * stage16, loB = loB0 + stage * stage16;
          for (int y = 0; y < KH; ++y) {
            const uint32_t acc = (i > 0 || y > 0) ? 1u : 0u;
            umma_bf16(tmem_base, umma_desc(hiD, loA), umma_desc(hiD, loB), idesc, acc);
            if (n_blocks > 1) umma_bf16(tmem_base + 192, umma_desc(hiD, loA), umma_desc(hiD, loB + xb16), idesc, acc);
            loA += 128;  // 16 rows x 128 B >> 4
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
      const int co = mt * 128 + q * 32 + lane;
      mbar_wait(tfull, 0, 32);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool valid = co < p.Cout;
#pragma unroll 1
      for (int b = 0; b < n_blocks; ++b) {
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          float* drow = p.dw + (long long)split * p.slab_stride + ((long long)co * 9 + r * 3 + s) * p.Cin + cg * 128 + b * 64;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_addr + b * 192 + s * 64 + hh * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 a = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                       __uint_as_float(v[i + 3]));
                __stcg(reinterpret_cast<float4*>(drow + hh * 32 + i), a);
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
