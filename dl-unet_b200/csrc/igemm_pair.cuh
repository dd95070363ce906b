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
//   * epilogue (16 warps of BOTH CTAs, on the CTA's own 128 TMEM lanes) + store warp; "accumulator drained" arrives on
//     the leader's TEMPTY[acc] (2 x 512 arrivals).
// Epilogue: SIXTEEN warps, warp = (TMEM lane quarter q, 16-column slice cs of the 64-column chunk).  The per-chunk chain
// tcgen05.ld -> bias / mask -> bf16 -> st.shared -> proxy fence -> arrive is latency bound (ncu on up-convolution 1,
// profiles/r2_ncu_upconv1_fwd_pair.txt: the eight warps of the first version issued 18 % of the time, the rest was
// fixed-latency waits, fetch bubbles behind 15 branches per chunk and ~350 instructions of index arithmetic per tile),
// and with K loops of 2-32 steps the epilogue IS the kernel.  Four warps per scheduler instead of two, forward / dgrad
// variants compiled separately (no mask, column-sum or bias code in the other one), unit coordinates by multiply-high.
// TMEM is allocated and freed with the cta_group::2 forms by the same warp of both CTAs, cluster barriers around.
//
// Stream-K (p.stream_k): the deep layers have few, long units (conv52c forward: 224 units of 144 K steps on 74 pairs =
// 3.03 waves, the last one 3 % full).  Instead of whole units round-robin, every pair takes one contiguous range of
// total_units * k_iters / n_pairs K STEPS of the unit sequence, so all pairs finish together.  A range that starts
// inside a unit computes that unit's remaining K steps FIRST and dumps the fp32 accumulator to a workspace slot
// (coalesced, [column][row]) and raises the slot's flag; a range that ends inside a unit computes its leading K steps
// LAST, and its epilogue adds the partial its successor dumped long before (a unit is split across at most two pairs:
// a pair's share is >= one unit, the host guarantees it).  The waiter has the HIGHER block index than the pair it waits
// for (ranges are handed out in reverse block order), so even with fewer resident CTAs than the grid the dumping pair
// is scheduled first: no circular wait.  Flags are reset by their consumer: 0 between launches, CUDA-graph safe.
// Deterministic: the split points are a function of the shape and the grid only.
#pragma once
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace ub {

// NSTAGE operand stages of 32 KB and NEPI output staging buffers of 16 KB (one 128-row x 64-column bf16 chunk each) share
// the 224 KB: <5, 4> for the convolutions (long K loops: the operand pipeline matters), <3, 8> for the up-convolutions,
// whose K loop is 2-32 steps and whose pace-maker is the drain of the TMA stores -- with NEPI buffers NEPI - 1 stores
// are in flight instead of one.
template <int NSTAGE_, int NEPI_>
struct IgemmPairCfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = 128 * 128;          // this CTA's half of the 256 columns
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int NEPI = NEPI_;
  static constexpr int EPI_BYTES = NEPI * 16384;
  static constexpr int BAR_BYTES = 2048;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert(NEPI >= 2 && NEPI <= 8 && NSTAGE >= 2 && NSTAGE <= 8, "ring sizes");
};

// One pair's sequence of work items: (unit, K-step range [kb, ke)).
struct PairItem {
  int u, kb, ke;
};
struct PairSched {
  int u, step, total_units, k_iters;
  long long g0, g1;   // stream-K: this pair's range of global K steps
  bool sk;
  __device__ __forceinline__ PairSched(int pair, int n_pairs, int total_units_, int k_iters_, bool sk_)
      : total_units(total_units_), k_iters(k_iters_), sk(sk_) {
    if (sk) {
      const int lp = n_pairs - 1 - pair;   // reverse order: the pair waited for has the lower block index
      const long long total = (long long)total_units * k_iters;
      g0 = total * lp / n_pairs;
      g1 = total * (lp + 1) / n_pairs;
      u = (int)(g0 / k_iters);
      step = 1;
    } else {
      g0 = g1 = 0;
      u = pair;
      step = n_pairs;
    }
  }
  __device__ __forceinline__ bool next(PairItem& it) {
    if (sk) {
      const long long base = (long long)u * k_iters;
      if (base >= g1) return false;
      it.u = u;
      it.kb = g0 > base ? (int)(g0 - base) : 0;
      it.ke = g1 - base < k_iters ? (int)(g1 - base) : k_iters;
    } else {
      if (u >= total_units) return false;
      it.u = u;
      it.kb = 0;
      it.ke = k_iters;
    }
    u += step;
    return true;
  }
};

constexpr int SK_SLOT_FLOATS = 2 * 256 * 128;   // one pair's dumped accumulators: [2 CTAs][256 columns][128 rows]

// bounded spin until *flag != 0 (acquire), then reset it for the next launch
__device__ __forceinline__ void sk_wait_and_reset(unsigned int* flag) {
  unsigned int v;
  const uint64_t t0 = globaltimer_ns();
  unsigned int spins = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v != 0) break;
    __nanosleep(100);
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > UB_WAIT_LIMIT_NS) {
      atomicExch(&g_device_error, 0x90000000u | (blockIdx.x & 0xffff));
      __threadfence_system();
      __trap();
    }
  }
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");
}

struct UnitCoord {
  int nt, x0, y0, img;
};
__device__ __forceinline__ UnitCoord decode_unit(const IgemmParams& p, int u, int rank) {
  UnitCoord c;
  const int mp = fast_div(u, p.mg_ntn);
  c.nt = u - mp * p.n_tiles_n;
  const int mt = mp * 2 + rank;
  const int r = fast_div(mt, p.mg_tx);
  c.x0 = (mt - r * p.tiles_x) * p.BW;
  const int it = fast_div(r, p.mg_ty);
  c.y0 = (r - it * p.tiles_y) * p.BH;
  c.img = it * p.BI;          // first image of the box
  return c;
}

constexpr int PAIR_EPI_WARPS = 16;
constexpr int PAIR_EPI_THREADS = PAIR_EPI_WARPS * 32;
constexpr int PAIR_THREADS = 128 + PAIR_EPI_THREADS;

template <int NSTAGE_, int NEPI_, bool DGRAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) igemm_pair_kernel(const __grid_constant__ IgemmParams p) {
  using C = IgemmPairCfg<NSTAGE_, NEPI_>;
  constexpr int BN = C::BN;
  constexpr int ET = PAIR_EPI_THREADS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* tfull = empty + C::NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint64_t* efull = tempty + 2;          // staging buffer written by the epilogue threads
  uint64_t* eempty = efull + C::NEPI;    // ... and read out by its TMA store
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(eempty + C::NEPI);
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
      mbar_init(&tempty[i], 2 * ET);    // the epilogue threads of both CTAs (only the leader's copy is used)
    }
    for (int i = 0; i < C::NEPI; ++i) {
      mbar_init(&efull[i], ET);
      mbar_init(&eempty[i], 1);
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
  const int m_tiles = p.img_tiles * p.tiles_y * p.tiles_x;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_units = m_pairs * p.n_tiles_n;
  const int k_iters = p.n_taps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    PairSched sched(pair, n_pairs, total_units, k_iters, p.stream_k != 0);
    PairItem item;
    while (sched.next(item)) {
      const UnitCoord uc = decode_unit(p, item.u, (int)rank);
      const int n0 = uc.nt * BN;
      if (DGRAD && p.mask != nullptr && item.kb == 0) {   // the pair that stores the unit reads the mask
        if (elect_one()) {
          for (int j = 0; j < BN / 64; ++j) tma_prefetch_l2_4d(&p.m_map, n0 + j * 64, uc.x0, uc.y0, uc.img);
        }
        __syncwarp();
      }
      int tap = item.kb / p.kchunks, kc = item.kb - tap * p.kchunks;
      for (int ki = item.kb; ki < item.ke; ++ki) {
        const CUtensorMap* am = &p.a_map[p.tap_map[tap]];
        const int ax = uc.x0 + p.tap_dx[tap], ay = uc.y0 + p.tap_dy[tap];
        mbar_wait(&empty[stage], phase ^ 1, 61);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full[stage], 2u * ((uint32_t)p.a_bytes + (uint32_t)C::B_BYTES));
          tma_load_4d_2sm(sa, am, &full[stage], kc * 64, ax, ay, uc.img);
          tma_load_2d_2sm(sb, &p.b_map, &full[stage], (tap * p.kchunks + kc) * 64, n0 + (int)rank * 128);
        }
        __syncwarp();
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
        if (++kc == p.kchunks) {
          kc = 0;
          ++tap;
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
    PairSched sched(pair, n_pairs, total_units, k_iters, p.stream_k != 0);
    PairItem item;
    for (; sched.next(item); ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1, 62);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int ki = item.kb; ki < item.ke; ++ki) {
        mbar_wait(&full[stage], phase, 63);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t so = stage * (C::STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(d_tmem, umma_desc(hiD, loA0 + so + 2 * k), umma_desc(hiD, loB0 + so + 2 * k), idesc,
                          (ki > item.kb || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty[stage]);
          if (ki == item.ke - 1) umma_commit_2sm(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------- TMA store warp (both CTAs): chunk c of the sequence lives in
    // staging buffer c % NEPI; up to NEPI - 1 stores are in flight, the oldest one's buffer is handed back as soon as it
    // has been read out
    uint32_t c = 0;
    PairSched sched(pair, n_pairs, total_units, k_iters, p.stream_k != 0);
    PairItem item;
    while (sched.next(item)) {
      if (item.kb > 0) continue;      // a dumped partial: nothing to store
      const UnitCoord uc = decode_unit(p, item.u, (int)rank);
      const int n0 = uc.nt * BN;
      for (int j = 0; j < BN / 64; ++j, ++c) {
        const uint32_t b = c % C::NEPI;
        mbar_wait(&efull[b], (c / C::NEPI) & 1, 66);
        if (lane == 0) {
          const int col0 = n0 + j * 64;
          const int mi = fast_div(col0, p.mg_omap);
          tma_store_4d(&p.o_map[mi], epi + b * 16384, col0 - mi * p.omap_div, uc.x0, uc.y0, uc.img);
          tma_store_commit();
          if (c >= (uint32_t)(C::NEPI - 2)) {
            tma_store_wait_read<C::NEPI - 2>();
            mbar_arrive(&eempty[(c - (C::NEPI - 2)) % C::NEPI]);
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
    if (DGRAD && p.dbias != nullptr) named_bar_arrive(6, ET + 32);   // the staging buffers are free: the epilogue reuses them
  } else if (warp >= 4) {
    // ----------------------------------------------------------- epilogue (both CTAs, own 128 TMEM lanes)
    const int q = warp & 3;                 // TMEM lane quarter = row quarter of the tile
    const int cs = (warp - 4) >> 2;         // 16-column slice of each 64-column chunk
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    // tile row -> (image of the box, y, x)
    const int bi = fast_div(row, p.mg_bwh), rr = row - bi * p.BW * p.BH;
    const int by = fast_div(rr, p.mg_bw), bx = rr - by * p.BW;
    const bool row_in_box = row < p.BW * p.BH * p.BI;
    const bool mask_on = DGRAD && p.mask != nullptr;
    // fused bias gradient (dgrad): per-thread running column sums in registers, one per 64-column chunk of the N <= 1024
    // outputs (static indices only), combined across the four row quarters in fixed order at the end (deterministic)
    const bool colsum_on = DGRAD && p.dbias != nullptr;
    float cacc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cacc[i] = 0.0f;
    int it = 0;
    uint32_t chunk_ctr = 0;
    PairSched sched(pair, n_pairs, total_units, k_iters, p.stream_k != 0);
    PairItem item;
    const int lp = n_pairs - 1 - pair;      // stream-K: this pair's position in the range order = its workspace slot
    for (; sched.next(item); ++it) {
      const UnitCoord uc = decode_unit(p, item.u, (int)rank);
      const int n0 = uc.nt * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (!DGRAD && p.bias != nullptr && (it == 0 || p.n_tiles_n > 1)) {
        if (it > 0) named_bar_sync(1, ET);
        if (et < BN) sbias[et] = __ldg(p.bias + (n0 + et) % p.bias_mod);
        named_bar_sync(1, ET);
      }
      mbar_wait(&tfull[acc], acc_phase, 64);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cs * 16;
      if (item.kb > 0) {
        // stream-K: the trailing K steps of a unit another pair finishes -- dump the accumulator, raise the flag
        float* slot = p.sk_ws + ((long long)lp * 2 + rank) * (SK_SLOT_FLOATS / 2) + cs * 16 * 128 + row;
#pragma unroll 1
        for (int j = 0; j < BN / 64; ++j) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + j * 64, v);
          tmem_ld_wait();
          float* d = slot + j * 64 * 128;
#pragma unroll
          for (int i = 0; i < 16; ++i) __stcg(d + i * 128, __uint_as_float(v[i]));
        }
        tc_fence_before();
        mbar_arrive_leader(&tempty[acc]);
        __threadfence();
        named_bar_sync(1, ET);
        if (et == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + lp * 2 + rank), "r"(1u) : "memory");
        continue;
      }
      const float* partial = nullptr;
      if (item.ke < k_iters) {
        // ... and the leading K steps of a unit whose rest the next range's pair dumped at its very start
        if (et == 0) sk_wait_and_reset(p.sk_flags + (lp + 1) * 2 + rank);
        named_bar_sync(1, ET);
        partial = p.sk_ws + ((long long)(lp + 1) * 2 + rank) * (SK_SLOT_FLOATS / 2) + cs * 16 * 128 + row;
      }
      const bool px_valid = row_in_box && (uc.img + bi) < p.n_img && (uc.x0 + bx) < p.out_W && (uc.y0 + by) < p.out_H;
      const __nv_bfloat16* mrow = nullptr;
      if (mask_on && px_valid)
        mrow = p.mask + (long long)(uc.img + bi) * p.mask_sn + (long long)(uc.y0 + by) * p.mask_sy +
               (long long)(uc.x0 + bx) * p.mask_sx + n0 + cs * 16;
#pragma unroll 1
      for (int j = 0; j < BN / 64; ++j, ++chunk_ctr) {
        const uint32_t eb = chunk_ctr % C::NEPI;
        uint8_t* ebuf = epi + eb * 16384;
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_addr + j * 64, v);
        uint4 m4[2];
        if (DGRAD && mrow != nullptr) {     // this pixel's 16 columns of the ReLU-backward mask source (32 B)
          m4[0] = __ldg(reinterpret_cast<const uint4*>(mrow + j * 64));
          m4[1] = __ldg(reinterpret_cast<const uint4*>(mrow + j * 64) + 1);
        }
        mbar_wait(&eempty[eb], ((chunk_ctr / C::NEPI) & 1) ^ 1, 65);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (partial != nullptr) {   // stream-K: the rest of this unit's K range, dumped as [column][128 rows]
          const float* pp = partial + j * 64 * 128;
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += __ldcg(pp + i * 128);
        }
        float csum = 0.0f;
        epi16_finish<!DGRAD>(f, (!DGRAD && p.bias != nullptr) ? sbias + j * 64 + cs * 16 : nullptr, p.relu, mask_on,
                             mrow != nullptr, m4, ebuf, row, cs, colsum_on ? &csum : nullptr, px_valid);
        if (DGRAD && colsum_on) {
          const int cidx = uc.nt * (BN / 64) + j;
#pragma unroll
          for (int i = 0; i < 16; ++i) cacc[i] += cidx == i ? csum : 0.0f;
        }
        fence_proxy_async_smem();
        mbar_arrive(&efull[eb]);
      }
      tc_fence_before();
      mbar_arrive_leader(&tempty[acc]);
    }
    if (colsum_on) {
      // one row of partial sums per CTA, no atomics: wait until the store warp has drained the staging buffers, then the
      // row quarters q = 0..3 add their register sums into them one after the other; rows of CTAs that do not exist
      // (grid smaller than the row count the host planned for) are zero-filled
      named_bar_sync(6, ET + 32);
      float* scol = reinterpret_cast<float*>(epi);
      const int N = p.n_tiles_n * BN;
      for (int qq = 0; qq < 4; ++qq) {
        if (q == qq && lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = i * 64 + cs * 16 + lane;
            if (col < N) scol[col] = (qq == 0 ? 0.0f : scol[col]) + cacc[i];
          }
        }
        named_bar_sync(1, ET);
      }
      const int len = N - p.dbias_c0;
      for (int ch = p.dbias_c0 + et; ch < N; ch += ET) {
        p.dbias[(long long)blockIdx.x * len + ch - p.dbias_c0] = scol[ch];
        for (int r = gridDim.x + blockIdx.x; r < p.dbias_rows; r += gridDim.x) p.dbias[(long long)r * len + ch - p.dbias_c0] = 0.0f;
      }
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
