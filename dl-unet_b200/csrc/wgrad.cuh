// Weight-gradient kernel of the deep 3x3 convolutions and of the 2x2 up-convolutions (reference: the
// autograd reverse of network.py:131-190, trainer.py:77).
//
//   conv    : dW[co, tap, ci]     = sum_{pixels p} dY[p, co]      * X[p + tap_offset, ci]
//   up-conv : dW[(a,b,co), ci]    = sum_{pixels p} dY[2p+(a,b), co] * X[p, ci]
//
// Both operands are NHWC, i.e. the reduction index (pixel) is the strided one and the channel is
// contiguous: both are "MN-major" UMMA operands.  A pixel chunk is a KW x KH patch of `rows` = KW*KH
// pixels (any multiple of 16 up to 128, chosen per layer so that the patches tile the map with the
// least padding); 64 channels of it are one TMA box = `rows` smem rows of 128 B (SWIZZLE_128B).
// M = 128 rows of the gradient = two such blocks LBO apart, N = BN input channels = BN/64 blocks.
//
// Work unit = (item, K split): item = (tap, 128-row tile, BN-column tile) for the convolutions -- the
// nine taps share dY, each reads X at its own offset -- and (128-row tile of the 4*Cout rows, column
// tile) for the up-convolutions, where the four (a,b) phases are rows of M read through four stride-2
// views of dY (Cout = 64: one tile holds two phases).  The kernel is PERSISTENT: CTA b runs units b,
// b + grid, ... (item fastest, so the CTAs resident at one time walk the same pixel chunks and share
// them in L2), with two TMEM accumulators so the read-out of unit i overlaps the MMAs of unit i+1.
//
// Deterministic split-K: every unit STORES its fp32 tile into its own slab, dw + split * slab_stride
// (no atomics); the slabs are summed in fixed order by the unpack kernel that also converts the
// packed layout to torch's (elementwise.cuh: unpack_all_kernel / unpack_grad_kernel).
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) WgradParams {
  CUtensorMap a_map[4];  // dY views (one per (a,b) phase for the up-convolution), box {64, KW, KH, 1}
  CUtensorMap b_map;     // X, box {64, KW, KH, 1}
  int n_taps;            // conv: 9 (taps are items); up-conv: 1 (the phases are rows of M)
  int tap_dx[9], tap_dy[9];
  int upconv;            // 1: M rows = (ab, co), 64-row block -> a_map[ab]
  int m_rows;            // Cout (conv) or 4 * Cout (up-conv)
  int m_tiles, n_tiles, items, splits, per;  // per = pixel chunks per split (no split is empty)
  int KW, KH, ksteps;    // ksteps = KW * KH / 16 MMAs per chunk
  int chunks_x, chunks_y, total_chunks;
  int Cout, Cin;
  int nstage, stage_bytes, blk_bytes;  // pipeline geometry, chosen by the host: blk_bytes = KW*KH*128
  int img_H;             // CONV1: rows per image of the 2-D bf16 image map in b_map
  int dbg;               // bring-up switches (ub_set_option 9): 1 no B loads, 2 no MMAs, 4 no TMEM read-out, 8 no init of constant rows
  float* dw;             // slabs [splits][packed layout]: conv [co][9][ci], up-conv [(ab,co)][ci]
  long long slab_stride; // elements between slabs
};

struct WgradCfg {
  static constexpr int MAX_STAGES = 20;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_LIMIT = 232448 - 1024 - BAR_BYTES;
};

struct WgradUnit {
  int split, tap, mt, nt;
};
__device__ __forceinline__ WgradUnit wgrad_unit(const WgradParams& p, int u) {
  WgradUnit w;
  w.split = u / p.items;
  int r = u - w.split * p.items;
  w.nt = r % p.n_tiles;
  r /= p.n_tiles;
  w.tap = r % p.n_taps;
  w.mt = r / p.n_taps;
  return w;
}

// CONV1 (BN = 32): the weight + bias gradient of conv11c (Cin = 1, network.py:131) on the same pipeline.
//   dW[co, tap] = sum_p dY[p, co] * x[p + tap],  db[co] = sum_p dY[p, co]
// is a GEMM with M = 64 channels, N = 32 columns (9 taps, a column of ones for the bias, the rest unused), K = pixels.
// A = dY as everywhere else (MN-major, one 64-channel block).  B is K-MAJOR, four 8-row SWIZZLE_128B groups of 64
// pixels (128 B) per row: group s = 0..2 receives ONE TMA box {64 pixels, 3 rows} of the bf16 image copy SHIFTED by s
// pixels (p.a_map[1 + s]: 2-D maps over three copies [n*H][pitch], copy_s[x] = image[x + s] -- a TMA box must start on
// a 16-byte boundary of global memory, which a one-pixel shift of a one-channel image is not), i.e. its rows 0..2 are
// the taps (r, s), r = 0..2; every box lands on a 1024-byte boundary of shared memory; row 0 of group 3 holds ones,
// written once per stage at kernel start.  Column of tap (r, s) = 8 s + r, bias = column 24.  A chunk is 64 pixels of one output row.
// Every unit stores one row of partial sums [640] = dW[64][9], db[64] into p.dw (summed in order by the unpack kernel).
//
// PAIR (BN = 256): the same kernel on CTA pairs (tcgen05 cta_group::2, see igemm_pair.cuh for the protocol): one MMA of
// M = 256 gradient rows x N = 256 columns per K step; each CTA loads its own 128 rows of dY and HALF of the X chunk
// (128 of the 256 input channels): 32 KB instead of 48 KB per CTA and chunk.  Needs m_rows % 256 == 0.
template <int BN, bool CONV1 = false, bool PAIR = false>
__device__ __forceinline__ void wgrad_kernel_impl(const WgradParams& p) {
  using C = WgradCfg;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // CTA (pair) index and count for the unit loop
  const int n_ctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int NB_B = PAIR ? BN / 128 : BN / 64;                        // 64-channel X blocks this CTA loads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NSTAGE = p.nstage, STAGE_BYTES = p.stage_bytes, BLK = p.blk_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  constexpr int ACC_COLS = CONV1 ? 32 : BN;        // TMEM columns between the two accumulators
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.items * p.splits;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 256 : 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (CONV1 && !(p.dbg & 8)) {
    // the rows of every stage's B tile that no TMA box writes: rows 3..7 of the tap groups and rows 1..7 of group 3 =
    // zeros (unused columns), row 0 of group 3 = ones (bias column).  All 64 values of a row are equal, so the 128-byte
    // swizzle is immaterial.
    for (int i = threadIdx.x; i < NSTAGE * 32 * 32; i += blockDim.x) {
      const int st = i / (32 * 32), r = (i / 32) % 32, c = i % 32;
      if (r >= 24 || (r & 7) >= 3)
        *reinterpret_cast<uint32_t*>(smem + st * STAGE_BYTES + BLK + r * 128 + c * 4) = r == 24 ? 0x3F803F80u : 0u;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer (whole warp loops, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int u = cta; u < total_units; u += n_ctas) {
      WgradUnit w = wgrad_unit(p, u);
      if (PAIR) w.mt = 2 * w.mt + (int)rank;
      // the two 64-row blocks of this M tile: tensor map + channel coordinate
      const int row_a = w.mt * 128, row_b = row_a + 64;
      const int nblk = row_b < p.m_rows ? 2 : 1;
      int map_a = 0, ch_a = row_a, map_b = 0, ch_b = row_b;
      if (p.upconv) {
        map_a = row_a / p.Cout;
        ch_a = row_a - map_a * p.Cout;
        map_b = (row_b / p.Cout) & 3;
        ch_b = row_b - (row_b / p.Cout) * p.Cout;
      }
      const CUtensorMap* am_a = &p.a_map[map_a];
      const CUtensorMap* am_b = &p.a_map[map_b];
      const int dx = p.tap_dx[w.tap], dy = p.tap_dy[w.tap];
      const uint32_t tx_bytes = CONV1  ? (uint32_t)BLK + ((p.dbg & 1) ? 0u : 3u * 384u)
                                : PAIR ? 2u * (uint32_t)(2 + NB_B) * (uint32_t)BLK      // both CTAs' loads, counted on the leader
                                       : (uint32_t)(nblk + BN / 64) * (uint32_t)BLK;
      const int c_begin = w.split * p.per;
      const int c_end = min(p.total_chunks, c_begin + p.per);
      int cx = c_begin % p.chunks_x;
      int r = c_begin / p.chunks_x;
      int cy = r % p.chunks_y;
      int img = r / p.chunks_y;
      for (int c = c_begin; c < c_end; ++c) {
        const int x0 = cx * p.KW, y0 = cy * p.KH;
        mbar_wait(&empty[stage], phase ^ 1, 5);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + (CONV1 ? BLK : 2 * BLK);
          if (PAIR) {
            if (leader) mbar_arrive_expect_tx(&full[stage], tx_bytes);
            tma_load_4d_2sm(sa, am_a, &full[stage], ch_a, x0, y0, img);
            tma_load_4d_2sm(sa + BLK, am_b, &full[stage], ch_b, x0, y0, img);
#pragma unroll
            for (int b = 0; b < NB_B; ++b)
              tma_load_4d_2sm(sb + b * BLK, &p.b_map, &full[stage], w.nt * BN + (int)rank * 128 + b * 64, x0 + dx, y0 + dy, img);
          } else {
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          tma_load_4d(sa, am_a, &full[stage], ch_a, x0, y0, img);
          if (CONV1) {
            if (!(p.dbg & 1))
#pragma unroll
            for (int sft = 0; sft < 3; ++sft)   // box {64 pixels, 3 rows} of the image copy shifted by sft: taps (0..2, sft)
              tma_load_2d(sb + sft * 1024, &p.a_map[1 + sft], &full[stage], x0 + ((p.dbg & 16) ? sft : 0), img * p.img_H + y0);
          } else {
            if (nblk > 1) tma_load_4d(sa + BLK, am_b, &full[stage], ch_b, x0, y0, img);
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              tma_load_4d(sb + b * BLK, &p.b_map, &full[stage], w.nt * BN + b * 64, x0 + dx, y0 + dy, img);
          }
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
        if (++cx == p.chunks_x) {
          cx = 0;
          if (++cy == p.chunks_y) {
            cy = 0;
            ++img;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // --------------------------------------------------------- MMA issuer (of the pair's leader)
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN, 1, CONV1 ? 0 : 1);
    const uint32_t hiD = umma_desc_hi(1024);
    // MN-major: LBO = distance between the 64-element M / N blocks, SBO = 1024 B between 8-row K groups;
    // CONV1: B is K-major (rows = taps, 128 B = 64 pixels of K per row): SBO = 1024 B between 8-row groups, 32 B per K step
    // CONV1: only 64 rows of M exist; LBO = 0 makes the second 64-row block an alias of the first (its accumulator rows
    // are never read), so a stage is 8 KB of dY + 4 KB of taps and 18 stages (144 KB of dY per SM) are in flight
    const uint32_t loA0 = umma_desc_lo(smem_u32(smem), CONV1 ? 0u : (uint32_t)BLK);
    const uint32_t loB0 = umma_desc_lo(smem_u32(smem) + (CONV1 ? BLK : 2 * BLK), CONV1 ? 16u : (uint32_t)BLK);
    constexpr uint32_t kstepB = CONV1 ? 2u : 128u;
    const uint32_t stage16 = (uint32_t)STAGE_BYTES >> 4;
    const int ksteps = p.ksteps;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int u = cta; u < total_units; u += n_ctas, ++it) {
      const int split = u / p.items;
      const int c_begin = split * p.per;
      const int n_iters = min(p.total_chunks, c_begin + p.per) - c_begin;
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1, 8);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int i = 0; i < n_iters; ++i) {
        mbar_wait(&full[stage], phase, 6);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t so = stage * stage16;
          // 16 pixels (K) per MMA = two 8-row groups = 2048 B (>>4 = 128)
#pragma unroll 4
          for (int k = 0; k < ((CONV1 && (p.dbg & 2)) ? 0 : ksteps); ++k) {
            if (PAIR)
              umma_bf16_2sm(d_tmem, umma_desc(hiD, loA0 + so + 128 * k), umma_desc(hiD, loB0 + so + kstepB * k), idesc,
                            (i > 0 || k > 0) ? 1u : 0u);
            else
              umma_bf16(d_tmem, umma_desc(hiD, loA0 + so + 128 * k), umma_desc(hiD, loB0 + so + kstepB * k), idesc,
                        (i > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) {
            umma_commit_2sm(&empty[stage]);
            if (i == n_iters - 1) umma_commit_2sm(&tfull[acc]);
          } else {
            umma_commit(&empty[stage]);
            if (i == n_iters - 1) umma_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------------- read-out: TMEM -> registers -> slab (plain stores)
    const int q = warp - 4;
    const int m = q * 32 + lane;  // accumulator row
    int it = 0;
    for (int u = cta; u < total_units; u += n_ctas, ++it) {
      WgradUnit w = wgrad_unit(p, u);
      if (PAIR) w.mt = 2 * w.mt + (int)rank;
      const int acc = it & 1;
      const int row0 = (w.mt * 2 + (m >> 6)) * 64;  // first row of this thread's 64-row block
      const bool valid = row0 < p.m_rows;
      // conv: packed row = co * 9 + tap;  up-conv: packed row = (ab * Cout + co) = the M row itself
      const long long prow = p.upconv ? (long long)(row0 + (m & 63)) : (long long)(row0 + (m & 63)) * p.n_taps + w.tap;
      float* drow = p.dw + (long long)w.split * p.slab_stride + prow * p.Cin + w.nt * BN;
      mbar_wait(&tfull[acc], (it >> 1) & 1, 7);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      if (CONV1 && !(p.dbg & 4)) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr, v);   // column 8 s + r = tap (r, s), column 24 = bias
        tmem_ld_wait();
        if (m < 64) {
          float* prow_ = p.dw + (long long)u * 640;
#pragma unroll
          for (int t = 0; t < 9; ++t) prow_[m * 9 + t] = __uint_as_float(v[(t % 3) * 8 + t / 3]);
          prow_[576 + m] = __uint_as_float(v[24]);
        }
      }
#pragma unroll 1
      for (int j = 0; j < (CONV1 ? 0 : BN / 32); ++j) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + j * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            __stcg(reinterpret_cast<float4*>(drow + j * 32 + i),
                   make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                               __uint_as_float(v[i + 3])));
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool CONV1 = false>
__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  wgrad_kernel_impl<BN, CONV1, false>(p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) wgrad_pair_kernel(const __grid_constant__ WgradParams p) {
  wgrad_kernel_impl<256, false, true>(p);
}

}  // namespace ub
