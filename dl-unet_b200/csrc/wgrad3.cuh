// 3x3 weight-gradient kernel, "all nine taps from one haloed patch" (supersedes the generic
// wgrad_kernel for the 3x3 convolutions; the up-convolutions keep using wgrad_kernel).
//
//   dW[co, (r,s), ci] = sum_{pixels (y,x)} dY[y, x, co] * X[y + r, x + s, ci]
//
// One CTA owns a (128-wide co tile, 32-wide ci tile, K split).  Per pixel chunk (16 x KH output
// pixels) it TMA-loads
//   * dY: box {64 co, 16, KH}   x (1 or 2 co blocks)  -> SWIZZLE_128B rows of 128 B, one per pixel
//   * X : box {32 ci, 18, KH+2}                       -> SWIZZLE_64B  rows of  64 B, one per pixel,
//     i.e. the chunk plus its 1-pixel halo, ONCE for all nine taps.
// Both operands are MN-major (the reduction index -- the pixel -- is the smem row).  A tap (r, s) is
// just a row offset ((y+r)*18 + x+s) into the X patch: UMMA applies the swizzle on absolute smem
// address bits (verified by ub_probe_run on B200), so a descriptor may start at any row.  The three
// s taps of one r are folded into ONE instruction by giving the B descriptor a leading-dimension
// offset of one row (64 B): N = 3 x 32 = 96 columns = (s, ci).  Three accumulators (r = 0,1,2) of 96
// fp32 columns live in TMEM; per 16-pixel K step the CTA issues 3 MMAs of 128 x 96 x 16.
//
// L2 -> SM traffic per chunk: (32 + 11.3) KB for 24 MMAs (1152 tensor cycles) ~= 38 B/clk/SM, versus
// ~190 B/clk for the tap-per-CTA kernel; dY is re-read Cin/32 times but by CTAs that run side by side
// (ci tile is the fastest grid index), i.e. from L2.
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) Wgrad3Params {
  CUtensorMap g_map;  // dY view, box {64, 16, KH, 1}, SWIZZLE_128B
  CUtensorMap x_map;  // X  view, box {32, 18, KH+2, 1}, SWIZZLE_64B
  int KH;
  int chunks_x, chunks_y, n_img;
  int m_tiles, ci_tiles, splits;
  int Cout, Cin;
  int fuse_s;  // 1: one N=96 MMA per r (LBO = 64 B);  0: three N=32 MMAs per r (diagnostic fallback)
  int nstage, stage_bytes, g_bytes;  // pipeline geometry (depends on KH), chosen by the host
  int per;                // pixel chunks per split (no split is empty)
  float* dw;              // slabs [splits][Cout][9][Cin] fp32: split s STORES its partial sums into slab s (deterministic;
                          // the slabs are summed in fixed order by the unpack kernel)
  long long slab_stride;  // elements between slabs
};

struct Wgrad3Cfg {
  static constexpr int MAX_STAGES = 8;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 232448 - 1024 - BAR_BYTES;
  static constexpr int TMEM_COLS = 512;                   // 3 x 96 = 288 used
  // stage = two co blocks of 16*KH pixels x 128 B  +  18*(KH+2) haloed pixels x 64 B (1 KB aligned)
  static int g_bytes(int kh) { return 2 * 16 * kh * 128; }
  static int stage_bytes(int kh) { return g_bytes(kh) + ((18 * (kh + 2) * 64 + 1023) / 1024) * 1024; }
  static int nstage(int kh) {
    const int n = SMEM_LIMIT / stage_bytes(kh);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
  static int smem_bytes(int kh) { return nstage(kh) * stage_bytes(kh) + BAR_BYTES + 1024; }
};

__global__ void __launch_bounds__(256, 1) wgrad3_kernel(const __grid_constant__ Wgrad3Params p) {
  using C = Wgrad3Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NSTAGE = p.nstage, STAGE_BYTES = p.stage_bytes, G_BYTES = p.g_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int w = blockIdx.x;
  const int ct = w % p.ci_tiles;  // fastest: neighbours share the dY chunk in L2
  w /= p.ci_tiles;
  const int mt = w % p.m_tiles;
  const int split = w / p.m_tiles;

  const int total_chunks = p.n_img * p.chunks_y * p.chunks_x;
  const int per = p.per;
  const int c_begin = split * per;
  const int c_end = min(total_chunks, c_begin + per);
  const int n_iters = c_end - c_begin;
  const int m_blocks = min(2, (p.Cout - mt * 128 + 63) / 64);
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
      {
        const uint32_t tx_bytes = (uint32_t)m_blocks * g_block_bytes + 18u * (KH + 2) * 64u;
        int stage = 0;
        uint32_t phase = 0;
        for (int c = c_begin; c < c_end; ++c) {
          const int cx = c % p.chunks_x;
          const int r2 = c / p.chunks_x;
          const int cy = r2 % p.chunks_y;
          const int img = r2 / p.chunks_y;
          const int x0 = cx * 16, y0 = cy * KH;
          mbar_wait(&empty[stage], phase ^ 1, 10);
          if (elect_one()) {
            uint8_t* sg = smem + stage * STAGE_BYTES;
            uint8_t* sx = sg + G_BYTES;
            mbar_arrive_expect_tx(&full[stage], tx_bytes);
            for (int b = 0; b < m_blocks; ++b)
              tma_load_4d(sg + b * g_block_bytes, &p.g_map, &full[stage], mt * 128 + b * 64, x0, y0, img);
            tma_load_4d(sx, &p.x_map, &full[stage], ct * 32, x0, y0, img);
          }
          __syncwarp();
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      {
        constexpr uint32_t idesc96 = umma_idesc_bf16(128, 96, 1, 1);
        constexpr uint32_t idesc32 = umma_idesc_bf16(128, 32, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_iters; ++i) {
          mbar_wait(&full[stage], phase, 11);
          tc_fence_after();
          const uint32_t sg = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sx = sg + G_BYTES;
          if (elect_one()) {
            // descriptors differ only in the 14-bit start-address field: build once, then add
            //   A: +2048 B per y (16 pixel rows of 128 B);  B: +18*64 B per patch row, tap r = +r rows
            uint64_t adesc = umma_smem_desc(sg, g_block_bytes, 1024, 0, 2);
            uint64_t bdesc = umma_smem_desc(sx, 64, 512, 0, 4);
            for (int y = 0; y < KH; ++y) {
              const uint32_t acc = (i > 0 || y > 0) ? 1u : 0u;
              if (p.fuse_s) {
                // B: 16 pixels (two 8-row groups, SBO = 512 B) x 96 = 3 blocks of 32 ci, block j = the
                // patch shifted by j pixels (LBO = one 64 B row): columns (s, ci)
                umma_bf16(tmem_base, adesc, bdesc, idesc96, acc);
                umma_bf16(tmem_base + 96, adesc, bdesc + 72, idesc96, acc);
                umma_bf16(tmem_base + 192, adesc, bdesc + 144, idesc96, acc);
              } else {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                  for (int s = 0; s < 3; ++s)
                    umma_bf16(tmem_base + r * 96 + s * 32, adesc, bdesc + r * 72 + s * 4, idesc32, acc);
              }
              adesc += 128;  // 2048 B >> 4
              bdesc += 72;   // 18 * 64 B >> 4
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
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      const int co = mt * 128 + q * 32 + lane;
      mbar_wait(tfull, 0, 12);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool valid = co < p.Cout;
#pragma unroll 1
      for (int t = 0; t < 9; ++t) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + t * 32, v);   // columns (r*3+s)*32 .. +31
        tmem_ld_wait();
        if (valid) {
          float* drow = p.dw + (long long)split * p.slab_stride + ((long long)co * 9 + t) * p.Cin + ct * 32;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace ub
