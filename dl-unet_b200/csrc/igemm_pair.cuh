// igemm.cuh's implicit GEMM on CTA PAIRS (tcgen05 cta_group::2): two CTAs of a cluster, on the two SMs of one TPC, run ONE
// tcgen05.mma of M = 256 pixels x N = 256 columns per K = 16 step.  Each CTA supplies its own 128-pixel A tile and only
// HALF of the weight slab (its 128 of the 256 columns), so the L2 -> SM operand traffic per unit of work drops from
// (16 + 32) KB to (16 + 16) KB per 2 x 128 x 256 x 64 MACs -- the deep layers are operand-bandwidth bound at N = 256
// (BN = 128 is 25-30 % slower than BN = 256 on every one of them, profiles/r2_bn_sweep_deep_layers.log), and their
// N is a multiple of 256.  Serves the same four cases as igemm_kernel (3x3 conv forward / dgrad, up-conv forward /
// dgrad) whenever N % 256 == 0.
//
// Protocol (leader = cluster rank 0):
//   * TMA producer (warp 0 of BOTH CTAs): waits its own EMPTY[stage]; the loads land in its own shared memory but count
//     their bytes on the LEADER's FULL[stage] (cta_group::2 loads with the peer bit of the barrier address cleared); the
//     leader arms FULL with the bytes of both CTAs.
//   * MMA issuer (warp 1 of the leader only): waits FULL[stage], issues four cta_group::2 MMAs, and commits with a
//     multicast arrive on EMPTY[stage] of both CTAs; the last commit of a tile multicasts on TFULL[acc] of both CTAs.
//   * epilogue + store warp (both CTAs): exactly igemm_kernel's, on the CTA's own 128 TMEM lanes; "accumulator drained"
//     arrives on the leader's TEMPTY[acc] (2 x 256 arrivals).
// TMEM is allocated and freed with the cta_group::2 forms by the same warp of both CTAs, cluster barriers around.
#pragma once
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace ub {

struct IgemmPairCfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = 128 * 128;          // this CTA's half of the 256 columns
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 2 * 16384;
  static constexpr int BAR_BYTES = 2048;
  static constexpr int SMEM_LIMIT = 232448 - 1024;
  static constexpr int NSTAGE = (SMEM_LIMIT - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;   // 6
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) igemm_pair_kernel(const __grid_constant__ IgemmParams p) {
  using C = IgemmPairCfg;
  constexpr int BN = C::BN;
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
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

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
      mbar_init(&tempty[i], 512);    // the epilogue threads of both CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs are initialised before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work units: (pair of adjacent M tiles, N tile); this CTA's M tile = 2 * pair tile + rank (may lie beyond the last
  // tile: its loads are zero-filled, its stores clipped)
  const int m_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_units = m_pairs * p.n_tiles_n;
  const int k_iters = p.n_taps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int u = pair; u < total_units; u += n_pairs) {
      const int nt = u % p.n_tiles_n;
      int mt = (u / p.n_tiles_n) * 2 + (int)rank;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
      if (p.mask != nullptr) {
        if (elect_one()) {
          for (int j = 0; j < BN / 64; ++j) tma_prefetch_l2_4d(&p.m_map, n0 + j * 64, x0, y0, img);
        }
        __syncwarp();
      }
      for (int tap = 0; tap < p.n_taps; ++tap) {
        const CUtensorMap* am = &p.a_map[p.tap_map[tap]];
        const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 61);
          if (elect_one()) {
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            uint8_t* sb = sa + C::A_BYTES;
            if (leader) mbar_arrive_expect_tx(&full[stage], 2u * ((uint32_t)p.a_bytes + (uint32_t)C::B_BYTES));
            tma_load_4d_2sm(sa, am, &full[stage], kc * 64, ax, ay, img);
            tma_load_2d_2sm(sb, &p.b_map, &full[stage], (tap * p.kchunks + kc) * 64, n0 + (int)rank * 128);
          }
          __syncwarp();
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // --------------------------------------------------------- MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
    const uint32_t hiD = umma_desc_hi(1024);
    const uint32_t loA0 = umma_desc_lo(smem_u32(smem), 16), loB0 = umma_desc_lo(smem_u32(smem) + C::A_BYTES, 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int u = pair; u < total_units; u += n_pairs, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1, 62);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int ki = 0; ki < k_iters; ++ki) {
        mbar_wait(&full[stage], phase, 63);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t so = stage * (C::STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(d_tmem, umma_desc(hiD, loA0 + so + 2 * k), umma_desc(hiD, loB0 + so + 2 * k), idesc,
                          (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty[stage]);
          if (ki == k_iters - 1) umma_commit_2sm(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------- TMA store warp (both CTAs; see igemm.cuh)
    uint32_t c = 0;
    for (int u = pair; u < total_units; u += n_pairs) {
      const int nt = u % p.n_tiles_n;
      int mt = (u / p.n_tiles_n) * 2 + (int)rank;
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
    // ----------------------------------------------------------- epilogue (both CTAs, own 128 TMEM lanes)
    const int q = warp & 3;
    const int hw = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const int by = row / p.BW, bx = row - by * p.BW;
    const bool row_in_box = row < p.BW * p.BH;
    const bool mask_on = p.mask != nullptr;
    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int u = pair; u < total_units; u += n_pairs, ++it) {
      const int nt = u % p.n_tiles_n;
      int mt = (u / p.n_tiles_n) * 2 + (int)rank;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (p.bias != nullptr && (it == 0 || p.n_tiles_n > 1)) {
        if (it > 0) named_bar_sync(1, 256);
        if (et < BN) sbias[et] = __ldg(p.bias + (n0 + et) % p.bias_mod);
        named_bar_sync(1, 256);
      }
      mbar_wait(&tfull[acc], acc_phase, 64);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const bool px_valid = row_in_box && img < p.n_img && (x0 + bx) < p.out_W && (y0 + by) < p.out_H;
      const __nv_bfloat16* mrow = nullptr;
      if (mask_on && px_valid)
        mrow = p.mask + (long long)img * p.mask_sn + (long long)(y0 + by) * p.mask_sy + (long long)(x0 + bx) * p.mask_sx +
               n0 + hw * 32;
#pragma unroll 1
      for (int j = 0; j < BN / 64; ++j, ++chunk_ctr) {
        uint8_t* ebuf = epi + (chunk_ctr & 1) * 16384;
        if (chunk_ctr >= 2) named_bar_sync(4 + (chunk_ctr & 1), 288);
        uint4 m4[4];
        if (mrow != nullptr) load_mask(mrow + j * 64, m4);
        epi_chunk(t_addr + j * 64 + hw * 32, p.bias != nullptr ? sbias + j * 64 + hw * 32 : nullptr, p.relu, mask_on,
                  mrow != nullptr, m4, ebuf, row, hw);
        fence_proxy_async_smem();
        named_bar_arrive(2 + (chunk_ctr & 1), 288);
      }
      tc_fence_before();
      mbar_arrive_leader(&tempty[acc]);
    }
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA frees TMEM or exits while its partner may still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace ub
