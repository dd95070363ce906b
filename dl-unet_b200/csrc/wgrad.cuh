// Weight-gradient kernel (reference: the autograd reverse of network.py:131-190,
// trainer.py:77).
//
//   dW[co, tap, ci] = sum_{pixels p} dY[p, co] * X[p + tap_offset, ci]
//
// Both operands are NHWC, i.e. the reduction index (pixel) is the strided one and
// the channel is contiguous: both are "MN-major" UMMA operands.  One 64-pixel
// patch (KW x KH) of 64 channels is one TMA box = one 8 KB SWIZZLE_128B block
// (64 rows of 128 B); M = 128 output channels = two such blocks LBO apart, N = BN
// input channels = BN/64 blocks.  One CTA owns one (tap, co tile, ci tile, K split)
// and reduces its share of pixel patches into one TMEM accumulator, then adds it
// to the fp32 packed gradient with vector atomics (split-K).
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) WgradParams {
  CUtensorMap a_map[4];  // dY views (per tap for the up-convolution), box {64, KW, KH, 1}
  CUtensorMap b_map;     // X, box {64, KW, KH, 1}
  int n_taps;
  int tap_amap[9], tap_dx[9], tap_dy[9];
  int m_tiles, n_tiles, splits;
  int KW, KH;
  int chunks_x, chunks_y, n_img;
  int Cout, Cin;
  int upconv_layout;  // 0: dw[(co*n_taps + tap)*Cin + ci]   1: dw[((tap*Cout) + co)*Cin + ci]
  float* dw;
};

template <int BN>
struct WgradCfg {
  static constexpr int A_BYTES = 2 * 8192;
  static constexpr int B_BYTES = (BN / 64) * 8192;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_LIMIT = 232448 - 1024;
  static constexpr int NSTAGE_RAW = (SMEM_LIMIT - BAR_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = WgradCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE_BYTES);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* tfull = empty + C::NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item decode: split fastest so the CTAs of one (tap, tile) run together
  int w = blockIdx.x;
  const int split = w % p.splits;
  w /= p.splits;
  const int nt = w % p.n_tiles;
  w /= p.n_tiles;
  const int mt = w % p.m_tiles;
  const int tap = w / p.m_tiles;

  const int total_chunks = p.n_img * p.chunks_y * p.chunks_x;
  const int per = (total_chunks + p.splits - 1) / p.splits;
  const int c_begin = split * per;
  const int c_end = min(total_chunks, c_begin + per);
  const int n_iters = c_end - c_begin;  // may be <= 0 for a trailing split

  const int m_valid_blocks = min(2, (p.Cout - mt * 128 + 63) / 64);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::NSTAGE; ++i) {
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
        const CUtensorMap* am = &p.a_map[p.tap_amap[tap]];
        const int dx = p.tap_dx[tap], dy = p.tap_dy[tap];
        const uint32_t tx_bytes = (uint32_t)(m_valid_blocks + BN / 64) * 8192u;
        int stage = 0;
        uint32_t phase = 0;
        for (int c = c_begin; c < c_end; ++c) {
          const int cx = c % p.chunks_x;
          int r = c / p.chunks_x;
          const int cy = r % p.chunks_y;
          const int img = r / p.chunks_y;
          const int x0 = cx * p.KW, y0 = cy * p.KH;
          mbar_wait(&empty[stage], phase ^ 1, 5);
          if (elect_one()) {
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            uint8_t* sb = sa + C::A_BYTES;
            mbar_arrive_expect_tx(&full[stage], tx_bytes);
            for (int b = 0; b < m_valid_blocks; ++b)
              tma_load_4d(sa + b * 8192, am, &full[stage], mt * 128 + b * 64, x0, y0, img);
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              tma_load_4d(sb + b * 8192, &p.b_map, &full[stage], nt * BN + b * 64, x0 + dx, y0 + dy, img);
          }
          __syncwarp();
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      {
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
        const uint32_t hiD = umma_desc_hi(1024);
        const uint32_t loA0 = umma_desc_lo(smem_u32(smem), 8192), loB0 = umma_desc_lo(smem_u32(smem) + C::A_BYTES, 8192);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_iters; ++i) {
          mbar_wait(&full[stage], phase, 6);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t so = stage * (C::STAGE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // 16 pixels (K) per MMA = two 8-row groups = 2048 B (>>4 = 128)
              umma_bf16(tmem_base, umma_desc(hiD, loA0 + so + 128 * k), umma_desc(hiD, loB0 + so + 128 * k), idesc,
                        (i > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (i == n_iters - 1) umma_commit(tfull);
          }
          __syncwarp();
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      const int co = mt * 128 + q * 32 + lane;
      mbar_wait(tfull, 0, 7);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool valid = co < p.Cout;
      float* drow = nullptr;
      if (valid) {
        const long long rowi = p.upconv_layout ? ((long long)tap * p.Cout + co) : ((long long)co * p.n_taps + tap);
        drow = p.dw + rowi * p.Cin + nt * BN;
      }
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + j * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 a = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                   __uint_as_float(v[i + 3]));
            atomicAdd(reinterpret_cast<float4*>(drow + j * 32 + i), a);
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
