// 3x3 convolution forward / data-gradient kernel for the wide, shallow layers (levels 1-2 of the
// U-Net): "haloed patch + shifted windows", optionally weight-stationary.
//
// igemm_kernel re-loads the 128-pixel activation tile once per tap (9 x 16 KB per 64 input channels)
// and the weight tile once per pixel tile; for N <= 128 output channels that is 110-190 B/clk/SM of
// L2->SM traffic against ~50 the chip sustains, so those layers run at 20-40 % of the tensor peak.
// Here one TMA box brings the (8*MT+2) x 18 pixel patch (output tile 8*MT x 16 plus the 1-pixel halo on
// each side, 64 channels, SWIZZLE_128B) and ALL NINE taps read it: tap (r, s) is a row offset
// (r*P + s)*128 B of the K-major A descriptor, the 8-row groups of the M = 128 tile are the 16 patch
// rows (SBO = P*128 B).  UMMA swizzles on absolute smem address bits (ub_probe_run), so this is legal.
// Weights: if all 9*K/64 weight slabs of the layer fit in shared memory next to two patches they are
// loaded ONCE per CTA (weight-stationary, persistent over tiles); otherwise they stream through a ring
// and each slab is used for MT = 2 adjacent pixel sub-tiles.
//   fwd  : taps read x[y0 + r, x0 + s]           (base offset 0)
//   dgrad: taps read dy[y0 + r - 2, x0 + s - 2]  (base offset -2, TMA zero fill = the padding), weights
//          pre-rotated by ub_pack_weight mode 1; optional fused ReLU-backward mask in the epilogue.
#pragma once
#include "elementwise.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

struct alignas(64) Igemm2Params {
  CUtensorMap a_map;  // input view, box {64, 8*MT+2, 18, 1}
  CUtensorMap b_map;  // packed weights [N][9*K], box {64, BN}
  CUtensorMap o_map;  // output view, box {64, 8, 16, 1}
  CUtensorMap m_map;  // ReLU-mask source view (same geometry as the output), used for L2 prefetch only
  int kchunks;
  int tiles_x, tiles_y, n_img, n_tiles_n;
  int base_off;  // 0 or -2
  int n_a, n_b, a_slot_bytes, stationary;
  const float* bias;
  int relu;
  const __nv_bfloat16* mask;
  long long mask_sn, mask_sy, mask_sx;
  int out_W, out_H;
  // fused 1x1 head (BN = 64 = all channels of the layer): logits / argmax from the fp32 accumulators
  const float* head_w;     // [2][64] fp32, nullptr = no head
  const float* head_b;     // [2]
  float* logits;           // [n][2][out_H][out_W] fp32 or nullptr
  uint8_t* argmax;         // [n][out_H][out_W] or nullptr (class 1 iff z1 > z0: ties -> 0, tester.py:30)
  int store_out;           // 0 = do not store the bf16 activation (inference; or: only its pooled form is wanted)
  // fused 2x2 max pool (POOL instantiation; network.py:133,139): the pooled tensor, the argmax codes and the skip
  // copy into the decoder's concat buffer are produced from the staged tile, see pool_fwd_kernel for the rules
  __nv_bfloat16* pool_y;   // dense [n][out_H/2][out_W/2][C]
  uint8_t* pool_idx;       // same shape, 1 byte per element
  __nv_bfloat16* pool_cat; // concat buffer view (dense in n, y, x) or nullptr
  int pool_C, cat_H, cat_W, cat_C, c_off;
  int reserved0;
  int dbias_rows;             // rows of the partial-sum buffer p.dbias (>= gridDim.x)
  int aux_bytes;              // bytes between the staging buffers and the mask tile buffers (2048, or 5120 with a head)
  float* dbias;               // fused column sums of the output (bias gradient of the consuming layer) or nullptr
  int dbias_c0;               // channels >= dbias_c0 are summed, into dbias[channel - dbias_c0]
  unsigned long long* prof;   // experiment (ub_set_prof_buffer): per-CTA clock sums of the epilogue phases of thread 128
};

constexpr int IG2_MAX_A = 4, IG2_MAX_B = 32;
// epilogue phase clocks (ub_set_prof_buffer, tests/prof_epi_clocks.py): compiled into the diagnostics build only -- the
// probe sums cost the sixteen-warp epilogue 14 registers of its 96
#ifdef UB_DIAGNOSTICS
constexpr bool IG2_PROF = true;
#else
constexpr bool IG2_PROF = false;
#endif

// The plain conv instantiations run SIXTEEN epilogue warps of 16 columns each (640 threads; see igemm_pair.cuh for why: the
// per-chunk epilogue chain is latency, and conv21c forward / conv11e dgrad were epilogue bound on eight warps); the fused
// head / pool instantiations keep eight warps of 32 columns.
template <bool HEAD, bool POOL>
constexpr int igemm2_threads() { return HEAD ? 384 : 640; }

// PAIR: the kernel runs on CTA pairs (launched with a cluster dimension of 2; tcgen05 cta_group::2, protocol as in
// igemm_pair.cuh).  The two CTAs work on two adjacent pixel tiles with the same output columns: ONE M = 256 MMA per step
// feeds both accumulators, each CTA keeps its own haloed patches and only HALF of every weight slab (its BN / 2 columns) in
// shared memory.  Why: the shallow layers are bound by shared-memory bandwidth -- an M = 128, N = 128 MMA reads 4 KB of A and
// 4 KB of B in 64 cycles, exactly the 128 B/clk of the SM, and the patch fills, staging writes and TMA-store reads come on
// top (N = 64: 6 KB per 54.5 cycles); with half of B per CTA the operand reads drop by 25 % (N = 128) / 17 % (N = 64), and
// the resident weights take half the space, which leaves room for deeper patch rings.
template <int BN, int MT, bool HEAD = false, bool POOL = false, bool PAIR = false>
__global__ void __launch_bounds__((igemm2_threads<HEAD, POOL>()), 1) igemm2_kernel(const __grid_constant__ Igemm2Params p) {
  constexpr int P = 8 * MT + 2;          // patch width in pixels
  constexpr bool WIDE = !HEAD && !POOL;
  // epilogue threads that drain the accumulators: sixteen warps of 16 columns (plain convs) or eight warps of 32 columns
  // (fused head; fused pool, whose warps 12..19 pool the staged chunks one step behind the drain)
  constexpr int ET = WIDE ? 512 : 256;
  constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 128;     // this CTA's part of one 64-channel weight slab
  constexpr int TMEM_COLS = 2 * MT * BN;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "tmem");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.n_a * p.a_slot_bytes;
  uint8_t* epi = sB + p.n_b * B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(epi + 2 * 16384);
  uint64_t* a_empty = a_full + IG2_MAX_A;
  uint64_t* b_full = a_empty + IG2_MAX_A;
  uint64_t* b_empty = b_full + IG2_MAX_B;
  uint64_t* tfull = b_empty + IG2_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(epi + 2 * 16384 + 1024);
  float* shead = reinterpret_cast<float*>(epi + 2 * 16384 + 2048);       // [2][64] weights + [2] bias (head only)
  float2* hpart = reinterpret_cast<float2*>(epi + 2 * 16384 + 2048 + 1024);  // [2 buffers][128 rows]
  uint64_t* m_full = reinterpret_cast<uint64_t*>(epi + 2 * 16384 + 640);     // ReLU-mask tiles (dgrad): 2 slots (after the 612 B above)
  uint64_t* m_empty = m_full + 2;
  uint8_t* mbuf = epi + 2 * 16384 + p.aux_bytes;                              // 2 x 16 KB, 1024-aligned
  float* scol = reinterpret_cast<float*>(epi + 2 * 16384 + p.aux_bytes - 1024); // [<= 256] column sums (dbias only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  // work units: PAIR -> (two adjacent pixel tiles, column tile), this CTA's pixel tile = 2 * pair tile + rank (may lie beyond
  // the last one: zero-filled loads, clipped stores); else one (pixel tile, column tile) per CTA
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int total_tiles = (PAIR ? (m_tiles + 1) >> 1 : m_tiles) * p.n_tiles_n;   // units

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    if ((!HEAD && !POOL) || p.store_out) tma_prefetch_desc(&p.o_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.n_a; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.n_b; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 2 * ET : ET);   // PAIR: both CTAs' epilogues arrive on the leader's copy
      mbar_init(&m_full[i], 1);
      mbar_init(&m_empty[i], ET);
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
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // barriers of both CTAs are initialised before anyone signals across
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = p.kchunks;
  const uint32_t a_bytes = (uint32_t)P * 18u * 128u;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform loop)
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    bool first = true;
    for (int t = unit0; t < total_tiles; t += unit_step) {
      const int nt = t % p.n_tiles_n;
      int mt = PAIR ? (t / p.n_tiles_n) * 2 + rank : t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int x0 = tx * 8 * MT + p.base_off, y0 = ty * 16 + p.base_off, n0 = nt * BN;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_empty[as], aph ^ 1, 20);
        if (elect_one()) {
          if (PAIR) {   // both CTAs' bytes are counted on the leader's barrier
            if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * a_bytes);
            tma_load_4d_2sm(sA + as * p.a_slot_bytes, &p.a_map, &a_full[as], kc * 64, x0, y0, img);
          } else {
            mbar_arrive_expect_tx(&a_full[as], a_bytes);
            tma_load_4d(sA + as * p.a_slot_bytes, &p.a_map, &a_full[as], kc * 64, x0, y0, img);
          }
        }
        __syncwarp();
        if (++as == p.n_a) {
          as = 0;
          aph ^= 1;
        }
        if (p.stationary) {
          if (first) {
            if (elect_one()) {
              for (int tap = 0; tap < 9; ++tap) {
                const int slot = kc * 9 + tap;
                if (PAIR) {
                  if (leader) mbar_arrive_expect_tx(&b_full[slot], 2u * (uint32_t)B_BYTES);
                  tma_load_2d_2sm(sB + slot * B_BYTES, &p.b_map, &b_full[slot], (tap * kchunks + kc) * 64, n0 + rank * (BN / 2));
                } else {
                  mbar_arrive_expect_tx(&b_full[slot], (uint32_t)B_BYTES);
                  tma_load_2d(sB + slot * B_BYTES, &p.b_map, &b_full[slot], (tap * kchunks + kc) * 64, n0);
                }
              }
            }
            __syncwarp();
          }
        } else {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1, 21);
            if (elect_one()) {
              if (PAIR) {
                if (leader) mbar_arrive_expect_tx(&b_full[bs], 2u * (uint32_t)B_BYTES);
                tma_load_2d_2sm(sB + bs * B_BYTES, &p.b_map, &b_full[bs], (tap * kchunks + kc) * 64, n0 + rank * (BN / 2));
              } else {
                mbar_arrive_expect_tx(&b_full[bs], (uint32_t)B_BYTES);
                tma_load_2d(sB + bs * B_BYTES, &p.b_map, &b_full[bs], (tap * kchunks + kc) * 64, n0);
              }
            }
            __syncwarp();
            if (++bs == p.n_b) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
      first = false;
    }
  } else if (warp == 1 && leader) {
    // --------------------------------------------------------------- MMA issuer (warp-uniform loop; PAIR: leader only)
    // Issue-side cost per MMA group must stay well below the tensor time of the group (4 x 54.5 clk for
    // N = 64): descriptors are (constant high word, low word = base + immediate), the tap loop is fully
    // unrolled, and a resident (stationary) weight slab is never waited on again.
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN, 0, 0);
    const uint32_t hiA = umma_desc_hi(P * 128), hiB = umma_desc_hi(1024);
    const uint32_t loA0 = umma_desc_lo(smem_u32(sA), 16), loB0 = umma_desc_lo(smem_u32(sB), 16);
    const uint32_t a_slot16 = (uint32_t)p.a_slot_bytes >> 4;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    bool resident = false;
    for (int t = unit0; t < total_tiles; t += unit_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1, 22);
      const uint32_t d0 = tmem_base + acc * MT * BN;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_full[as], aph, 23);
        const uint32_t loA = loA0 + as * a_slot16;
        if (p.stationary) {
          if (!resident)
            for (int tap = 0; tap < 9; ++tap) mbar_wait(&b_full[kc * 9 + tap], 0, 24);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t loB = loB0 + kc * 9 * (B_BYTES >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t aoff = (((tap / 3) * P + (tap % 3) + 8 * sub) * 128 + k * 32) >> 4;
                  const uint32_t boff = (tap * B_BYTES + k * 32) >> 4;
                  if (PAIR)
                    umma_bf16_2sm(d0 + sub * BN, umma_desc(hiA, loA + aoff), umma_desc(hiB, loB + boff), idesc,
                                  (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  else
                    umma_bf16(d0 + sub * BN, umma_desc(hiA, loA + aoff), umma_desc(hiB, loB + boff), idesc,
                              (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            if (PAIR) {
              umma_commit_2sm(&a_empty[as]);
              if (kc == kchunks - 1) umma_commit_2sm(&tfull[acc]);
            } else {
              umma_commit(&a_empty[as]);
              if (kc == kchunks - 1) umma_commit(&tfull[acc]);
            }
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[bs], bph, 24);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t loB = loB0 + bs * (B_BYTES >> 4);
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t aoff = (((tap / 3) * P + (tap % 3) + 8 * sub) * 128 + k * 32) >> 4;
                  if (PAIR)
                    umma_bf16_2sm(d0 + sub * BN, umma_desc(hiA, loA + aoff), umma_desc(hiB, loB + 2 * k), idesc,
                                  (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  else
                    umma_bf16(d0 + sub * BN, umma_desc(hiA, loA + aoff), umma_desc(hiB, loB + 2 * k), idesc,
                              (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
              }
              if (PAIR) {
                umma_commit_2sm(&b_empty[bs]);
                if (tap == 8) {
                  umma_commit_2sm(&a_empty[as]);
                  if (kc == kchunks - 1) umma_commit_2sm(&tfull[acc]);
                }
              } else {
                umma_commit(&b_empty[bs]);
                if (tap == 8) {
                  umma_commit(&a_empty[as]);
                  if (kc == kchunks - 1) umma_commit(&tfull[acc]);
                }
              }
            }
            __syncwarp();
            if (++bs == p.n_b) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        if (++as == p.n_a) {
          as = 0;
          aph ^= 1;
        }
      }
      resident = true;
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ ReLU-mask tile loader (dgrad)
    // The mask of a chunk (the 8 x 16 pixel x 64 channel tile of the layer's input activation) comes through TMA
    // into one of two 16 KB slots, two chunks ahead of its use: per-thread global loads of 64 B at a 128 B stride
    // cost the epilogue ~2200 cycles per chunk (32 sectors per request, L2 latency on the critical path).
    if (p.mask != nullptr) {
      uint32_t c = 0;
      for (int t = unit0; t < total_tiles; t += unit_step) {
        const int nt = t % p.n_tiles_n;
        int mt = PAIR ? (t / p.n_tiles_n) * 2 + rank : t / p.n_tiles_n;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int img = mt / p.tiles_y;
        for (int sub = 0; sub < MT; ++sub) {
          for (int j = 0; j < BN / 64; ++j, ++c) {
            const uint32_t sl = c & 1;
            mbar_wait(&m_empty[sl], ((c >> 1) & 1) ^ 1, 26);
            if (elect_one()) {
              mbar_arrive_expect_tx(&m_full[sl], 16384u);
              tma_load_4d(mbuf + sl * 16384, &p.m_map, &m_full[sl], nt * BN + j * 64, (tx * MT + sub) * 8, ty * 16, img);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 3 && !HEAD && !POOL) {
    // ------------------------------------------------------------------ TMA store warp (plain convs)
    // Issuing a TMA store blocks the issuing thread for ~600 cycles and an epilogue warp that has to do it holds
    // up the chunk barrier of all eight (measured with the clock probes below: 2400 cycles per chunk against
    // ~800 of actual epilogue work).  Here the epilogue warps only ARRIVE on FULL[b] when staging buffer b is
    // written and go on with the next chunk; this warp waits for FULL[b], issues the store, and arrives on
    // EMPTY[b'] once the store that read buffer b' has finished with it.  Named barriers 2,3 = FULL, 4,5 = EMPTY.
    uint32_t c = 0;
    for (int t = unit0; t < total_tiles; t += unit_step) {
      const int nt = t % p.n_tiles_n;
      int mt = PAIR ? (t / p.n_tiles_n) * 2 + rank : t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      for (int sub = 0; sub < MT; ++sub) {
        for (int j = 0; j < BN / 64; ++j, ++c) {
          const uint32_t b = c & 1;
          named_bar_sync(2 + b, ET + 32);
          // hand the OTHER buffer back first (its store was issued a whole chunk ago and has long read it): the
          // ~600-cycle issue of this chunk's store must not sit between the epilogue and its next buffer.  (Issuing
          // first and waiting for the previous store only -- two stores in flight -- measured the same.)
          if (c >= 1) {
            if (elect_one()) tma_store_wait_read<0>();
            __syncwarp();
            named_bar_arrive(4 + ((c - 1) & 1), ET + 32);
          }
          if (elect_one()) {
            tma_store_4d(&p.o_map, epi + b * 16384, nt * BN + j * 64, (tx * MT + sub) * 8, ty * 16, img);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  } else if (POOL && warp >= 12) {
    // ------------------------------------------------------------------------------------ pool stage (fused pool only)
    // Eight warps, one chunk behind the drain warps: wait for FULL[b] (named barriers 2,3; the drain warps arrive), one
    // thread issues the TMA store of the staged conv output, all 256 pool the staged 8 x 16 pixel tile 2x2 -- thread =
    // (pooled pixel, channel octet) -- and write pooled values, argmax codes and the skip half of the concat buffer
    // straight to global memory; EMPTY[b] (barriers 4,5) once the tile has been pooled and read by its store.
    const int pt = threadIdx.x - 384;
    const int pp = pt >> 3, c8 = pt & 7;
    const int py = pp >> 2, pxl = pp & 3;
    const int Ho = p.out_H >> 1, Wo = p.out_W >> 1;
    uint32_t c = 0;
    for (int t = unit0; t < total_tiles; t += unit_step) {
      const int nt = t % p.n_tiles_n;
      int mt = PAIR ? (t / p.n_tiles_n) * 2 + rank : t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int y0 = ty * 16, n0 = nt * BN;
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const int x0 = (tx * MT + sub) * 8;
        const int oy = (y0 >> 1) + py, ox = (x0 >> 1) + pxl;
#pragma unroll 1
        for (int j = 0; j < BN / 64; ++j, ++c) {
          const uint32_t b = c & 1;
          const uint8_t* ebuf = epi + b * 16384;
          named_bar_sync(2 + b, 512);   // FULL[b]
          if (pt == 0 && p.store_out) {
            tma_store_4d(&p.o_map, ebuf, n0 + j * 64, x0, y0, img);
            tma_store_commit();
          }
          if (oy < Ho && ox < Wo && img < p.n_img) {
            const int r00 = (2 * py) * 8 + 2 * pxl;
            float v[4][8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = r00 + (k >> 1) * 8 + (k & 1);
              unpack8(lds_u4(ebuf + r * 128 + ((c8 ^ (r & 7)) << 4)), v[k]);
            }
            float m[8];
            uint32_t code[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              m[ch] = v[0][ch];
              code[ch] = 0;
#pragma unroll
              for (int k = 1; k < 4; ++k) {
                const float tv = v[k][ch];
                if (tv > m[ch] || tv != tv) {   // first maximum in window order, NaN wins (ATen)
                  m[ch] = tv;
                  code[ch] = k;
                }
              }
            }
            const uint4 o = pack8(m);
            const int ch = n0 + j * 64 + c8 * 8;
            const long long pix = ((long long)img * Ho + oy) * Wo + ox;
            *reinterpret_cast<uint4*>(p.pool_y + pix * p.pool_C + ch) = o;
            uint2 cc;
            cc.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
            cc.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
            *reinterpret_cast<uint2*>(p.pool_idx + pix * p.pool_C + ch) = cc;
            if (p.pool_cat != nullptr) {
              const int cy = oy - p.c_off, cx = ox - p.c_off;
              if (cy >= 0 && cy < p.cat_H && cx >= 0 && cx < p.cat_W)
                *reinterpret_cast<uint4*>(p.pool_cat + (((long long)img * p.cat_H + cy) * p.cat_W + cx) * p.cat_C + ch) = o;
            }
          }
          if (pt == 0 && p.store_out) tma_store_wait_read<0>();
          named_bar_arrive(4 + b, 512);   // EMPTY[b]
        }
      }
    }
    if (pt == 0) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------ epilogue
    // 8 warps: warp w reads TMEM lanes 32*(w%4).. and owns the 32-column half h = (w-4)/4 of every chunk
    // (16 warps, WIDE: the 16-column slice h = (w-4)/4 of 0..3)
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const int by = row >> 3, bx = row & 7;
    const bool mask_on = p.mask != nullptr;
    constexpr bool head_on = HEAD && BN == 64;   // separate instantiation: the plain conv epilogue stays lean
    const bool store_out = (!head_on && !POOL) || p.store_out != 0;
    if (head_on) {
      if (et < 128) shead[et] = __ldg(p.head_w + et);
      if (et < 2) shead[128 + et] = __ldg(p.head_b + et);
      named_bar_sync(1, 256);
    }
    // fused bias gradient (dgrad, sixteen-warp epilogue, one column tile: N = BN <= 128 -- the host does not ask for it
    // otherwise): every thread keeps running sums of the values it stores, one register per column of its slice of each
    // 64-column chunk (static indices: the chunk loop is unrolled); they are reduced over the warp's 32 pixel rows once,
    // after the last tile, and combined across the four row quarters in fixed order (deterministic, no atomics)
    const bool colsum_on = WIDE && p.dbias != nullptr;
    float crun[WIDE ? (BN / 64) * 16 : 1];
#pragma unroll
    for (int i = 0; i < (WIDE ? (BN / 64) * 16 : 1); ++i) crun[i] = 0.0f;
    int it = 0;
    uint32_t chunk_ctr = 0;
    long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
    const long long prof_t0 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
    for (int t = unit0; t < total_tiles; t += unit_step, ++it) {
      const int nt = t % p.n_tiles_n;
      int mt = PAIR ? (t / p.n_tiles_n) * 2 + rank : t / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int y0 = ty * 16, n0 = nt * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (p.bias != nullptr && (it == 0 || p.n_tiles_n > 1)) {
        if (it > 0) named_bar_sync(1, ET);   // every warp has finished reading the previous tile's slice
        if (et < BN) sbias[et] = __ldg(p.bias + n0 + et);
        named_bar_sync(1, ET);
      }
      const long long c0 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
      mbar_wait(&tfull[acc], acc_phase, 25);
      tc_fence_after();
      if (IG2_PROF && p.prof && et == 0) prof_acc[0] += clock64() - c0;
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const int x0 = (tx * MT + sub) * 8;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + sub) * BN;
        const bool px_valid = (x0 + bx) < p.out_W && (y0 + by) < p.out_H && img < p.n_img;
        // two staging buffers; plain convs hand them to the store warp (FULL / EMPTY named barriers), the fused
        // head / pool variants keep one barrier per chunk among the epilogue warps (see igemm.cuh)
        if constexpr (WIDE) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j, ++chunk_ctr) {
            uint8_t* ebuf = epi + (chunk_ctr & 1) * 16384;
            const long long c1 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_addr + j * 64 + h * 16, v);
            if (chunk_ctr >= 2) named_bar_sync(4 + (chunk_ctr & 1), ET + 32);   // EMPTY[b]
            const long long c2 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            uint4 mcur[2];
            if (mask_on) {   // this chunk's mask tile: same rows / swizzle as the staging buffer
              const uint32_t sl = chunk_ctr & 1;
              mbar_wait(&m_full[sl], (chunk_ctr >> 1) & 1, 27);
              const uint8_t* mrow = mbuf + sl * 16384 + row * 128;
#pragma unroll
              for (int i = 0; i < 2; ++i) mcur[i] = lds_u4(mrow + (((h * 2 + i) ^ (row & 7)) << 4));
              mbar_arrive(&m_empty[sl]);
            }
            tmem_ld_wait();
            const long long c3 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (p.bias != nullptr || p.relu)
              epi16_finish<true>(f, p.bias != nullptr ? sbias + j * 64 + h * 16 : nullptr, p.relu, false, true, mcur, ebuf, row, h);
            else
              epi16_finish<false>(f, nullptr, 0, mask_on, true, mcur, ebuf, row, h, nullptr, true, colsum_on ? crun + j * 16 : nullptr);
            const long long c4 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            fence_proxy_async_smem();
            named_bar_arrive(2 + (chunk_ctr & 1), ET + 32);   // FULL[b]: the store warp takes it from here
            if (IG2_PROF && p.prof && et == 0) {
              const long long c5 = clock64();
              prof_acc[1] += c5 - c1;          // the whole chunk
              prof_acc[2] += c2 - c1;          // TMEM load issue + wait for the staging buffer (EMPTY)
              prof_acc[3] += c3 - c2;          // mask tile wait + read, TMEM load wait
              prof_acc[4] += c4 - c3;          // arithmetic, staging stores, column sums
              prof_acc[5] += 1;
            }
          }
        }
#pragma unroll 1
        for (int j = 0; !WIDE && j < BN / 64; ++j, ++chunk_ctr) {
          uint8_t* ebuf = epi + (chunk_ctr & 1) * 16384;
          if (POOL && chunk_ctr >= 2) named_bar_sync(4 + (chunk_ctr & 1), 512);   // EMPTY[b]: pooled and read by its store
          uint4 mcur[4];
          if (mask_on) {   // this chunk's mask tile: same rows / swizzle as the staging buffer
            const uint32_t sl = chunk_ctr & 1;
            mbar_wait(&m_full[sl], (chunk_ctr >> 1) & 1, 27);
            const uint8_t* mrow = mbuf + sl * 16384 + row * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) mcur[i] = lds_u4(mrow + (((h * 4 + i) ^ (row & 7)) << 4));
            mbar_arrive(&m_empty[sl]);
          }
          float2 hp = make_float2(0.0f, 0.0f);
          const long long c1 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
          epi_chunk(t_addr + j * 64 + h * 32, p.bias != nullptr ? sbias + j * 64 + h * 32 : nullptr, p.relu, mask_on,
                    true, mcur, ebuf, row, h, head_on ? shead + h * 32 : nullptr, &hp, store_out || POOL);
          const long long c2 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
          if (head_on && h == 1) hpart[(chunk_ctr & 1) * 128 + row] = hp;
          fence_proxy_async_smem();
          if constexpr (POOL) {
            named_bar_arrive(2 + (chunk_ctr & 1), 512);   // FULL[b]: the pool stage takes it from here
            if (IG2_PROF && p.prof && et == 0) {
              prof_acc[1] += c2 - c1;          // TMEM load + arithmetic + staging
              prof_acc[5] += 1;
            }
          } else {
            if (et == 0 && store_out) tma_store_wait_read<0>();
            const long long c3 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            named_bar_sync(1, 256);
            const long long c4 = (IG2_PROF && p.prof && et == 0) ? clock64() : 0;
            if (et == 0 && store_out) {
              tma_store_4d(&p.o_map, ebuf, n0 + j * 64, x0, y0, img);
              tma_store_commit();
            }
            if (IG2_PROF && p.prof && et == 0) {
              prof_acc[1] += c2 - c1;          // TMEM load + arithmetic + staging
              prof_acc[2] += c3 - c2;          // fence + wait for the previous TMA store's smem reads
              prof_acc[3] += c4 - c3;          // chunk barrier (skew between the 8 epilogue warps)
              prof_acc[4] += clock64() - c4;   // TMA store issue
              prof_acc[5] += 1;
            }
          }
          if (head_on && h == 0 && px_valid) {
            // the other half's partial sums were written before the barrier; this buffer is rewritten two
            // chunks from now, i.e. after every reader has passed the next barrier
            const float2 o = hpart[(chunk_ctr & 1) * 128 + row];
            const float z0 = hp.x + o.x + shead[128], z1 = hp.y + o.y + shead[129];
            const long long pix = (long long)(y0 + by) * p.out_W + (x0 + bx);
            const long long plane = (long long)p.out_H * p.out_W;
            if (p.logits != nullptr) {
              p.logits[(long long)img * 2 * plane + pix] = z0;
              p.logits[((long long)img * 2 + 1) * plane + pix] = z1;
            }
            if (p.argmax != nullptr) p.argmax[(long long)img * plane + pix] = (z1 > z0) ? 1 : 0;
          }
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_leader(&tempty[acc]);
      else mbar_arrive(&tempty[acc]);
    }
    if (et == 0 && HEAD) tma_store_wait_all<0>();
    if (colsum_on) {
      // one flush per CTA, no atomics: the row quarters q = 0..3 add their register sums into scol one after the
      // other, then the CTA stores its row of the partial sums [dbias_rows][N - c0]; CTAs beyond the grid the host
      // planned for (ub_set_sm_limit) do not exist, so their rows are zero-filled here
      const int N = BN;
      float cfin[BN / 64];     // lanes l and l ^ 16: the warp's total of column h * 16 + (l & 15) of chunk j
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = crun[WIDE ? j * 16 + i : 0];
        cfin[j] = colsum16_reduce(t);
      }
      for (int qq = 0; qq < 4; ++qq) {
        if (q == qq && lane < 16) {
          const int col = h * 16 + lane;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) scol[j * 64 + col] = (qq == 0 ? 0.0f : scol[j * 64 + col]) + cfin[j];
        }
        named_bar_sync(1, ET);
      }
      const int ch = et, len = N - p.dbias_c0;
      if (ch < N && ch >= p.dbias_c0) {
        p.dbias[(long long)blockIdx.x * len + ch - p.dbias_c0] = scol[ch];
        for (int r = gridDim.x + blockIdx.x; r < p.dbias_rows; r += gridDim.x) p.dbias[(long long)r * len + ch - p.dbias_c0] = 0.0f;
      }
    }
    if (IG2_PROF && p.prof && et == 0) {
      for (int i = 0; i < 6; ++i) p.prof[blockIdx.x * 8 + i] = (unsigned long long)prof_acc[i];
      p.prof[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - prof_t0);
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();   // no CTA frees TMEM or exits while its partner may still signal it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace ub
