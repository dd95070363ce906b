// Descriptor-semantics probes (test infrastructure, not on the product path).
//
// Question they answer on real hardware: can a UMMA shared-memory descriptor start at
// an arbitrary 128-byte row of a SWIZZLE_128B tile that TMA wrote (i.e. is the swizzle
// a function of the absolute smem address), and can the 8-row groups be an arbitrary
// multiple of 128 B apart (SBO)?  If yes, one haloed activation patch in smem can feed
// all nine taps of a 3x3 convolution ("shifted window"), cutting L2->SM traffic ~9x.
#pragma once
#include "ptx.cuh"

namespace ub {

struct alignas(64) ProbeParams {
  CUtensorMap a_map;  // [320 rows][64] bf16, box {64, 64}
  CUtensorMap b_map;  // [64 rows][64] bf16, box {64, 64}
  int mode;           // 0: K-major A shifted/strided;  1: MN-major, B shifted by s rows
  int shift, pitch, use_base_offset;
  float* out;  // [128][64]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;               // 320 rows * 128 B = 40 KB
  uint8_t* sb = smem + 320 * 128;   // 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 8192);
  uint64_t* mma_bar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 320 * 128 + 8192);
    for (int i = 0; i < 5; ++i) tma_load_2d(sa + i * 8192, &p.a_map, bar, 0, i * 64);
    tma_load_2d(sb, &p.b_map, bar, 0, 0);
    mbar_wait(bar, 0, 8);
    tc_fence_after();
    if (p.mode == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t a0 = smem_u32(sa) + p.shift * 128;
      const uint32_t bo = p.use_base_offset ? ((a0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_smem_desc(a0 + k * 32, 16, p.pitch * 128, bo);
        const uint64_t bd = umma_smem_desc(smem_u32(sb) + k * 32, 16, 1024);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
    } else {
      // D[m][n] = sum_k A[k][m] * B[k + shift][n]; A = rows 0..63 of the a tile viewed MN-major
      // (only 64 valid M columns: M rows 64..127 read the next 8 KB block = rows 64..127 as "block 1")
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t b0 = smem_u32(sa) + 128 * 128 + p.shift * 128;  // B operand = rows 128+shift.. of the a tile
      const uint32_t bo = p.use_base_offset ? ((b0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_smem_desc(smem_u32(sa) + k * 2048, 8192, 1024);
        const uint64_t bd = umma_smem_desc(b0 + k * 2048, 8192, 1024, bo);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
    }
    umma_commit(mma_bar);
    mbar_wait(mma_bar, 0, 9);
  }
  __syncthreads();
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
    for (int j = 0; j < 2; ++j) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + j * 32, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) p.out[row * 64 + j * 32 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

// MMA issue-rate probe: `iters` x 4 back-to-back tcgen05.mma on operands that already sit in shared
// memory (contents irrelevant), one commit at the end; reports SM cycles.  Separates the tensor
// pipe + smem operand fetch rate from everything TMA / L2 related.
struct MmaRateParams {
  int m;            // UMMA M (64 or 128)
  int n;            // UMMA N
  int a_mn, b_mn;   // 0 = K-major, 1 = MN-major
  int b_layout;     // 2 = SW128, 4 = SW64
  int b_lbo, b_sbo; // bytes
  int iters;
  long long* out;   // [2]: cycles, n_mma
};

__global__ void __launch_bounds__(128, 1) probe_mma_rate_kernel(const MmaRateParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;            // 64 KB
  uint8_t* sb = smem + 65536;    // 64 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(p.m, p.n, p.a_mn, p.b_mn);
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = p.a_mn ? umma_smem_desc(smem_u32(sa) + k * 2048, 8192, 1024, 0, 2)
                                   : umma_smem_desc(smem_u32(sa) + k * 32, 16, 1024, 0, 2);
        const uint64_t bd = p.b_mn ? umma_smem_desc(smem_u32(sb) + k * (p.b_layout == 2 ? 2048 : 1024), p.b_lbo, p.b_sbo, 0, p.b_layout)
                                   : umma_smem_desc(smem_u32(sb) + k * 32, 16, 1024, 0, 2);
        umma_bf16(tmem + (k & 1) * 256, ad, bd, idesc, 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0, 13);
    const long long t1 = clock64();
    p.out[0] = t1 - t0;
    p.out[1] = (long long)p.iters * 4;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace ub
