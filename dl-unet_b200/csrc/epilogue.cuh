// Shared epilogue step of the conv kernels: one thread's 32 columns of one 128-row x 64-column chunk.
//   TMEM -> registers -> (+bias from smem) -> ReLU -> bf16 -> ReLU-backward mask -> swizzled smem row.
// The 128 B rows of the staging buffer use the SWIZZLE_128B pattern the TMA store expects.
#pragma once
#include "ptx.cuh"

namespace ub {

// Optional fused 1x1 head (finalconv, network.py:190): `shead` = this thread's 32-column slice of the two head
// weight rows in smem ([2][64] floats, row pitch 64); the partial dot products of the fp32, post-ReLU, NOT yet
// bf16-rounded activations are returned in *head_part.  `stage` = false skips the bf16 staging write (inference:
// the activation itself is never stored).
// The ReLU-backward mask row (64 B: this thread's pixel, its 32 columns of the source activation).
__device__ __forceinline__ void load_mask(const __nv_bfloat16* mcols, uint4 (&m4)[4]) {
  const uint4* mp = reinterpret_cast<const uint4*>(mcols);
#pragma unroll
  for (int i = 0; i < 4; ++i) m4[i] = __ldg(mp + i);
}

// f[32]: this thread's pixel, 32 fp32 columns fresh from the accumulator -> bias, ReLU, optional head partials,
// bf16, optional ReLU-backward mask, swizzled staging row `row` (128 B rows, SWIZZLE_128B as the TMA store expects).
__device__ __forceinline__ void epi_finish(float (&f)[32], const float* sbias, int relu, bool mask_on, bool mask_valid,
                                           const uint4 (&m4)[4], uint8_t* ebuf, int row, int h,
                                           const float* shead = nullptr, float2* head_part = nullptr,
                                           bool stage = true, float* col_acc = nullptr, bool sum_valid = true) {
  if (sbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = lds_f4(sbias + i);  // same address in every lane: broadcast
      f[i] += b4.x;
      f[i + 1] += b4.y;
      f[i + 2] += b4.z;
      f[i + 3] += b4.w;
    }
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
  }
  if (shead != nullptr) {
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 w0 = lds_f4(shead + i);       // broadcast reads
      const float4 w1 = lds_f4(shead + 64 + i);
      s0 = fmaf(f[i], w0.x, s0); s0 = fmaf(f[i + 1], w0.y, s0); s0 = fmaf(f[i + 2], w0.z, s0); s0 = fmaf(f[i + 3], w0.w, s0);
      s1 = fmaf(f[i], w1.x, s1); s1 = fmaf(f[i + 1], w1.y, s1); s1 = fmaf(f[i + 2], w1.z, s1); s1 = fmaf(f[i + 3], w1.w, s1);
    }
    *head_part = make_float2(s0, s1);
  }
  if (!stage) return;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o;
    o.x = pack_bf16x2(f[c * 8 + 0], f[c * 8 + 1]);
    o.y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
    o.z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]);
    o.w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
    if (mask_on) {
      // ReLU backward on packed bf16 pairs: a bf16 is > 0 exactly when its bits, read as int16, are > 0
      if (mask_valid) {
        o.x &= __vcmpgts2(m4[c].x, 0u);
        o.y &= __vcmpgts2(m4[c].y, 0u);
        o.z &= __vcmpgts2(m4[c].z, 0u);
        o.w &= __vcmpgts2(m4[c].w, 0u);
      } else {
        o = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    const int chunk = (h * 4 + c) ^ (row & 7);
    sts_u4(ebuf + row * 128 + chunk * 16, o);
    if (col_acc != nullptr) {   // keep the stored (masked, bf16-rounded) values for the column sums below
      f[c * 8 + 0] = __uint_as_float(o.x << 16); f[c * 8 + 1] = __uint_as_float(o.x & 0xffff0000u);
      f[c * 8 + 2] = __uint_as_float(o.y << 16); f[c * 8 + 3] = __uint_as_float(o.y & 0xffff0000u);
      f[c * 8 + 4] = __uint_as_float(o.z << 16); f[c * 8 + 5] = __uint_as_float(o.z & 0xffff0000u);
      f[c * 8 + 6] = __uint_as_float(o.w << 16); f[c * 8 + 7] = __uint_as_float(o.w & 0xffff0000u);
    }
  }
  if (col_acc != nullptr) {
    if (!sum_valid) {   // a row that is no pixel of the tensor (patch smaller than 128 rows, tile beyond the last one)
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = 0.0f;
    }
    // Column sums over the warp's 32 pixel rows (the bias gradient of the layer that consumes this gradient):
    // recursive halving -- each step a lane keeps one half of its columns and receives the partner's sums of that
    // half -- 31 shuffles instead of 160, and lane l ends up with the total of column l of this warp's 32-column
    // slice, returned in *col_acc (the caller keeps per-thread running sums in registers: a fixed summation order,
    // so the bias gradients are bit-reproducible).
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float keep = up ? f[i + s] : f[i];
        const float send = up ? f[i] : f[i + s];
        f[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    *col_acc = f[0];
  }
}


// mask_on: apply the ReLU-backward mask; mask_valid = false zeroes the row (pixel outside the image), else m4 holds
// the mask row, loaded by the caller (igemm2 requests it one chunk ahead so that its L2 latency is off the
// tfull -> TMA-store critical path).
__device__ __forceinline__ void epi_chunk(uint32_t taddr, const float* sbias, int relu, bool mask_on, bool mask_valid,
                                          const uint4 (&m4)[4], uint8_t* ebuf, int row, int h,
                                          const float* shead = nullptr, float2* head_part = nullptr,
                                          bool stage = true, float* col_acc = nullptr, bool sum_valid = true,
                                          const float* partial = nullptr) {
  uint32_t v[32];
  tmem_ld_32x32b_x32(taddr, v);
  tmem_ld_wait();
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (partial != nullptr) {   // stream-K: the rest of this unit's K range, dumped by another CTA as [column][128 rows]
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] += __ldcg(partial + i * 128);
  }
  epi_finish(f, sbias, relu, mask_on, mask_valid, m4, ebuf, row, h, shead, head_part, stage, col_acc, sum_valid);
}

// ---- sixteen-warp epilogues: one thread's 16 columns (slice cs of the 64-column chunk) of one 128-row chunk.
// FWD: + bias (smem, broadcast reads), ReLU.  Otherwise (dgrad): ReLU-backward mask m4 (this pixel's 16 columns of the
// mask source, 32 B; mask_valid = false zeroes the row) and, with col_acc, the sums of the STORED values over the warp's
// 32 pixel rows: recursive halving over lane bits 3..0 (15 shuffles: a lane keeps one half of its columns and receives the
// partner's sums of that half), then the two half-warps are added -- lanes l and l ^ 16 both end with the total of column
// l & 15 of this slice (sum_valid = false: the row is no pixel of the tensor and counts as zero).
// run16 (instead of col_acc): the stored values are added to the thread's OWN sixteen running sums (its pixel row, its
// 16 columns) -- no shuffles per chunk; the caller reduces them over the warp's rows once per kernel (colsum16_reduce).
__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t m) {   // 0xffff per half that is > 0 (one HSET2.BF16.BM)
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&m), __floats2bfloat162_rn(0.0f, 0.0f));
}

// sums of sixteen per-thread values over the warp's 32 lanes by recursive halving: lanes l and l ^ 16 return the total
// of f[l & 15] (15 + 1 shuffles)
__device__ __forceinline__ float colsum16_reduce(float (&f)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? f[i + s] : f[i];
      const float send = up ? f[i] : f[i + s];
      f[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return f[0] + __shfl_xor_sync(0xffffffffu, f[0], 16);
}

template <bool FWD>
__device__ __forceinline__ void epi16_finish(float (&f)[16], const float* sbias16, int relu, bool mask_on, bool mask_valid,
                                             const uint4 (&m4)[2], uint8_t* ebuf, int row, int cs, float* col_acc = nullptr,
                                             bool sum_valid = true, float* run16 = nullptr) {
  if (FWD) {
    if (sbias16 != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b4 = lds_f4(sbias16 + i);   // same address in every lane: broadcast
        f[i] += b4.x;
        f[i + 1] += b4.y;
        f[i + 2] += b4.z;
        f[i + 3] += b4.w;
      }
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
    }
  }
  uint4 o[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    o[c].x = pack_bf16x2(f[c * 8 + 0], f[c * 8 + 1]);
    o[c].y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
    o[c].z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]);
    o[c].w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
    if (!FWD && mask_on) {
      // ReLU backward on packed bf16 pairs
      if (mask_valid) {
        o[c].x &= bf16x2_gt0_mask(m4[c].x);
        o[c].y &= bf16x2_gt0_mask(m4[c].y);
        o[c].z &= bf16x2_gt0_mask(m4[c].z);
        o[c].w &= bf16x2_gt0_mask(m4[c].w);
      } else {
        o[c] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    // 128 B staging rows in the SWIZZLE_128B pattern the TMA store expects
    sts_u4(ebuf + row * 128 + (((cs * 2 + c) ^ (row & 7)) << 4), o[c]);
  }
  if (!FWD && run16 != nullptr) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 oc = sum_valid ? o[c] : make_uint4(0u, 0u, 0u, 0u);
      run16[c * 8 + 0] += __uint_as_float(oc.x << 16); run16[c * 8 + 1] += __uint_as_float(oc.x & 0xffff0000u);
      run16[c * 8 + 2] += __uint_as_float(oc.y << 16); run16[c * 8 + 3] += __uint_as_float(oc.y & 0xffff0000u);
      run16[c * 8 + 4] += __uint_as_float(oc.z << 16); run16[c * 8 + 5] += __uint_as_float(oc.z & 0xffff0000u);
      run16[c * 8 + 6] += __uint_as_float(oc.w << 16); run16[c * 8 + 7] += __uint_as_float(oc.w & 0xffff0000u);
    }
  }
  if (!FWD && col_acc != nullptr) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 oc = sum_valid ? o[c] : make_uint4(0u, 0u, 0u, 0u);
      f[c * 8 + 0] = __uint_as_float(oc.x << 16); f[c * 8 + 1] = __uint_as_float(oc.x & 0xffff0000u);
      f[c * 8 + 2] = __uint_as_float(oc.y << 16); f[c * 8 + 3] = __uint_as_float(oc.y & 0xffff0000u);
      f[c * 8 + 4] = __uint_as_float(oc.z << 16); f[c * 8 + 5] = __uint_as_float(oc.z & 0xffff0000u);
      f[c * 8 + 6] = __uint_as_float(oc.w << 16); f[c * 8 + 7] = __uint_as_float(oc.w & 0xffff0000u);
    }
    *col_acc = colsum16_reduce(f);
  }
}

}  // namespace ub
