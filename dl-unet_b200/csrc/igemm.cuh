// Implicit-GEMM kernel for the tensor-pipe half of the U-Net path.
//
//   D[pixel, n] = sum_{tap} sum_{k < 64*kchunks} A_tap[pixel, k] * B[n, tap*64*kchunks + k]
//
// "pixel" runs over a BW x BH spatial patch of one image (<= 128 pixels = the
// UMMA M dimension), A_tap is an NHWC bf16 activation tensor read through a 4-D
// TMA tensor map at a per-tap (dx, dy) offset, B is a K-major packed bf16 weight
// matrix.  The same kernel therefore serves (reference: network.py:131-190)
//   * the 3x3 valid convolutions, forward        (9 taps, offsets (s, r))
//   * their data gradients                        (9 taps, offsets (s-2, r-2), TMA
//                                                  out-of-bounds zero fill is the padding)
//   * the 2x2 stride-2 up-convolutions, forward   (1 tap, N = 4*Cout, the four (a,b)
//                                                  column groups are stored through four
//                                                  stride-2 output tensor maps)
//   * their data gradients                        (4 taps through four stride-2 input maps)
//
// Pipeline: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM
// allocator, warps 4..7 = epilogue (TMEM -> registers -> bias/ReLU/mask -> bf16 ->
// swizzled smem -> TMA store).  Two TMEM accumulators so the epilogue of tile i
// overlaps the MMAs of tile i+1; persistent over tiles.
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int IG_MAX_TAPS = 9;

struct alignas(64) IgemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap o_map[4];
  int n_taps, kchunks;
  int tiles_x, tiles_y, n_img, n_tiles_n;
  int BW, BH;
  int a_bytes;  // BW*BH*128: bytes one A box deposits
  int tap_map[IG_MAX_TAPS];
  int tap_dx[IG_MAX_TAPS];
  int tap_dy[IG_MAX_TAPS];
  // epilogue
  const float* bias;  // may be null
  int bias_mod;       // bias index = column % bias_mod
  int relu;
  const __nv_bfloat16* mask;  // may be null: ReLU-backward mask source (value > 0 passes)
  long long mask_sn, mask_sy, mask_sx;  // element strides of the mask tensor (channel stride 1)
  int out_W, out_H;                     // valid output extent (for the mask read predicate)
  int omap_div;                         // column c -> o_map[c / omap_div], channel coord c % omap_div
};

template <int BN>
struct IgemmCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 2 * 16384;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_LIMIT = 232448 - 1024;
  static constexpr int NSTAGE_RAW = (SMEM_LIMIT - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "tmem cols");
};

template <int BN>
__global__ void __launch_bounds__(256, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  using C = IgemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* tfull = empty + C::NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      tma_prefetch_desc(&p.a_map[i]);
      tma_prefetch_desc(&p.o_map[i]);
    }
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
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

  const int m_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int total_tiles = m_tiles * p.n_tiles_n;
  const int k_iters = p.n_taps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer
    // The whole warp runs the loop (all values stay warp-uniform, so ptxas keeps descriptors and
    // coordinates in uniform registers); one elected lane issues the copies.
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles_n;
        int mt = t / p.n_tiles_n;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int img = mt / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const CUtensorMap* am = &p.a_map[p.tap_map[tap]];
          const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1, 1);
            if (elect_one()) {
              uint8_t* sa = smem + stage * C::STAGE_BYTES;
              uint8_t* sb = sa + C::A_BYTES;
              mbar_arrive_expect_tx(&full[stage], (uint32_t)p.a_bytes + (uint32_t)C::B_BYTES);
              tma_load_4d(sa, am, &full[stage], kc * 64, ax, ay, img);
              tma_load_2d(sb, &p.b_map, &full[stage], (tap * p.kchunks + kc) * 64, n0);
            }
            __syncwarp();
            if (++stage == C::NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t hiD = umma_desc_hi(1024);
      const uint32_t loA0 = umma_desc_lo(smem_u32(smem), 16), loB0 = umma_desc_lo(smem_u32(smem) + C::A_BYTES, 16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t so = stage * (C::STAGE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // +32 B per K=16 step inside the 128 B swizzle row (encoded >>4 -> +2)
              umma_bf16(d_tmem, umma_desc(hiD, loA0 + so + 2 * k), umma_desc(hiD, loB0 + so + 2 * k), idesc,
                        (ki > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (ki == k_iters - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------------- epilogue
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const int by = row / p.BW, bx = row - by * p.BW;
    const bool row_in_box = row < p.BW * p.BH;
    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int nt = t % p.n_tiles_n;
      int mt = t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const bool px_valid = row_in_box && (x0 + bx) < p.out_W && (y0 + by) < p.out_H;
      const __nv_bfloat16* mrow = nullptr;
      if (p.mask != nullptr && px_valid)
        mrow = p.mask + (long long)img * p.mask_sn + (long long)(y0 + by) * p.mask_sy + (long long)(x0 + bx) * p.mask_sx;
#pragma unroll 1
      for (int j = 0; j < BN / 64; ++j, ++chunk_ctr) {
        uint8_t* ebuf = epi + (chunk_ctr & 1) * 16384;
        // the TMA store that last read this buffer (two chunks ago) must be done reading
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        const int col0 = n0 + j * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_addr + j * 64 + h * 32, v);
          tmem_ld_wait();
          const int cbase = col0 + h * 32;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
            const float* bp = p.bias + (cbase % p.bias_mod);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
              f[i] += b4.x;
              f[i + 1] += b4.y;
              f[i + 2] += b4.z;
              f[i + 3] += b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (p.mask != nullptr) {
            if (mrow != nullptr) {
              const uint4* mp = reinterpret_cast<const uint4*>(mrow + cbase);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 m4 = __ldg(mp + i);
                const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  const uint32_t lo = mw[w] & 0xffffu, hi = mw[w] >> 16;
                  if (!(lo != 0 && lo < 0x8000u)) f[i * 8 + w * 2] = 0.0f;
                  if (!(hi != 0 && hi < 0x8000u)) f[i * 8 + w * 2 + 1] = 0.0f;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = 0.0f;
            }
          }
          // 32 channels = 64 B = four 16 B chunks of this row's 128 B swizzle line
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 o;
            o.x = pack_bf16x2(f[c * 8 + 0], f[c * 8 + 1]);
            o.y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
            o.z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]);
            o.w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
            const int chunk = (h * 4 + c) ^ (row & 7);
            *reinterpret_cast<uint4*>(ebuf + row * 128 + chunk * 16) = o;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          const int mi = col0 / p.omap_div;
          const int cc = col0 - mi * p.omap_div;
          tma_store_4d(&p.o_map[mi], ebuf, cc, x0, y0, img);
          tma_store_commit();
        }
      }
      // all of this thread's TMEM reads of accumulator `acc` are complete
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace ub
