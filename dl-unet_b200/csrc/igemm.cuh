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
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

constexpr int IG_MAX_TAPS = 9;

struct alignas(64) IgemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap o_map[4];
  CUtensorMap m_map;  // ReLU-mask source view, box {64, BW, BH, BI}: L2 prefetch only
  int n_taps, kchunks;
  int tiles_x, tiles_y, n_img, n_tiles_n;
  int BW, BH;
  // igemm_pair_kernel, 3x3 convolutions: the box also spans BI images (BW * BH * BI <= 128 pixel rows, row = (image, y, x)),
  // so that small maps fill the 128 MMA rows (28 x 28: 4 x 4 x 8 = 128 rows instead of 28 x 4 = 112); img_tiles =
  // ceil(n_img / BI).  Everything else: BI = 1, img_tiles = n_img.
  int BI, img_tiles;
  int a_bytes;  // BW*BH*BI*128: bytes one A box deposits
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
  int dbg;                              // experiment switch (ub_set_option 5): 1 = do not wait for earlier TMA stores
  // fused column sums of the stored output (igemm_pair_kernel; igemm_kernel with one column tile) (bias gradient of the layer that produced the input of
  // this dgrad): row blockIdx.x of dbias[dbias_rows][N - dbias_c0], channels >= dbias_c0; nullptr = off
  float* dbias;
  int dbias_c0, dbias_rows;
  // igemm_pair_kernel only: stream-K (see igemm_pair.cuh).  sk_ws: [pairs][2 CTAs][256 columns][128 rows] fp32 partial
  // accumulators, sk_flags: [pairs][2] "partial ready" flags (0 between launches)
  int stream_k;
  float* sk_ws;
  unsigned int* sk_flags;
  // igemm_pair_kernel only: fast_div magics of n_tiles_n, tiles_x, tiles_y, BW, omap_div (filled by the launcher)
  uint32_t mg_ntn, mg_tx, mg_ty, mg_bw, mg_bwh, mg_omap;   // mg_bwh: BW * BH
};

template <int BN>
struct IgemmCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 2 * 16384;
  static constexpr int BAR_BYTES = 2048;  // mbarriers, TMEM slot, 256 floats of per-tile bias
  static constexpr int SMEM_LIMIT = 232448 - 1024;
  static constexpr int NSTAGE_RAW = (SMEM_LIMIT - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "tmem cols");
};

template <int BN>
__global__ void __launch_bounds__(384, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  using C = IgemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* tfull = empty + C::NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(epi + C::EPI_BYTES + 1024);

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
      mbar_init(&tempty[i], 256);
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
        if (p.mask != nullptr) {
          // the epilogue reads the ReLU-mask rows of this tile later: pull them into L2 now
          if (elect_one()) {
            for (int j = 0; j < BN / 64; ++j) tma_prefetch_l2_4d(&p.m_map, n0 + j * 64, x0, y0, img);
          }
          __syncwarp();
        }
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
  } else if (warp == 3) {
    // ------------------------------------------------------- TMA store warp (see igemm2.cuh for the measurements)
    // The epilogue warps ARRIVE on FULL[b] when staging buffer b is written and carry on; this warp waits for FULL[b],
    // hands the other buffer back (EMPTY) once its store has read it, and issues the store (~600 cycles of issue
    // that used to sit in an epilogue warp in front of the chunk barrier).  Named barriers 2,3 = FULL, 4,5 = EMPTY.
    uint32_t c = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t % p.n_tiles_n;
      int mt = t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
      for (int j = 0; j < BN / 64; ++j, ++c) {
        const uint32_t b = c & 1;
        named_bar_sync(2 + b, 288);
        if (c >= 1) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
          named_bar_arrive(4 + ((c - 1) & 1), 288);
        }
        if (elect_one()) {
          const int col0 = n0 + j * 64;
          const int mi = col0 / p.omap_div;
          tma_store_4d(&p.o_map[mi], epi + b * 16384, col0 - mi * p.omap_div, x0, y0, img);
          tma_store_commit();
        }
        __syncwarp();
      }
    }
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  } else if (warp >= 4) {
    // ----------------------------------------------------------- epilogue
    // 8 warps: warp w may read TMEM lanes 32*(w%4).. and owns the 32-column half hw = (w-4)/4 of every chunk
    const int q = warp & 3;
    const int hw = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const int by = row / p.BW, bx = row - by * p.BW;
    const bool row_in_box = row < p.BW * p.BH;
    const bool mask_on = p.mask != nullptr;
    // fused bias gradient (dgrad flavour, one column tile: N = BN; the host asks for it only then): lane l keeps the running
    // total of column hw * 32 + l of each 64-column chunk over the rows of its quarter; fixed order, no atomics
    const bool colsum_on = p.dbias != nullptr;
    float cacc[BN / 64];
#pragma unroll
    for (int i = 0; i < BN / 64; ++i) cacc[i] = 0.0f;
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
      if (p.bias != nullptr && (it == 0 || p.n_tiles_n > 1)) {
        // per-tile bias slice -> smem (every read of the previous slice happened before the last chunk barrier)
        if (it > 0) named_bar_sync(1, 256);   // every warp has finished reading the previous tile's slice
        if (et < BN) sbias[et] = __ldg(p.bias + (n0 + et) % p.bias_mod);
        named_bar_sync(1, 256);
      }
      mbar_wait(&tfull[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const bool px_valid = row_in_box && (x0 + bx) < p.out_W && (y0 + by) < p.out_H;
      const __nv_bfloat16* mrow = nullptr;
      if (mask_on && px_valid)
        mrow = p.mask + (long long)img * p.mask_sn + (long long)(y0 + by) * p.mask_sy + (long long)(x0 + bx) * p.mask_sx +
               n0 + hw * 32;
#pragma unroll 1
      for (int j = 0; j < BN / 64; ++j, ++chunk_ctr) {
        // Two staging buffers handed to the store warp: wait until the store that last read this one is done with
        // it (EMPTY), write it, make the writes visible to the async proxy, arrive on FULL.
        uint8_t* ebuf = epi + (chunk_ctr & 1) * 16384;
        if (chunk_ctr >= 2) named_bar_sync(4 + (chunk_ctr & 1), 288);
        uint4 m4[4];
        // the mask row is requested first: its latency overlaps the TMEM load
        if (mrow != nullptr) load_mask(mrow + j * 64, m4);
        float csum = 0.0f;
        epi_chunk(t_addr + j * 64 + hw * 32, p.bias != nullptr ? sbias + j * 64 + hw * 32 : nullptr, p.relu, mask_on,
                  mrow != nullptr, m4, ebuf, row, hw, nullptr, nullptr, true, colsum_on ? &csum : nullptr, px_valid);
        if (colsum_on) {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) cacc[i] += (j == i) ? csum : 0.0f;
        }
        fence_proxy_async_smem();
        named_bar_arrive(2 + (chunk_ctr & 1), 288);
      }
      // all of this thread's TMEM reads of accumulator `acc` are complete
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
    if (colsum_on) {
      // one row of partial sums per CTA: the row quarters q = 0..3 add their sums one after the other into the (unused:
      // no bias in the dgrad flavour) bias slice, then the CTA stores row blockIdx.x of dbias[dbias_rows][N - c0]; rows of
      // CTAs that do not exist (grid smaller than the row count the host planned for) are zero-filled
      float* scol = sbias;
      for (int qq = 0; qq < 4; ++qq) {
        if (q == qq) {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) {
            const int col = i * 64 + hw * 32 + lane;
            scol[col] = (qq == 0 ? 0.0f : scol[col]) + cacc[i];
          }
        }
        named_bar_sync(1, 256);
      }
      const int len = BN - p.dbias_c0;
      for (int ch = p.dbias_c0 + et; ch < BN; ch += 256) {
        p.dbias[(long long)blockIdx.x * len + ch - p.dbias_c0] = scol[ch];
        for (int r = gridDim.x + blockIdx.x; r < p.dbias_rows; r += gridDim.x) p.dbias[(long long)r * len + ch - p.dbias_c0] = 0.0f;
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
