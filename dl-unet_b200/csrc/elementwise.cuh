// HBM-bound kernels of the U-Net path: first conv (Cin = 1), 2x2 max-pool with
// argmax codes (+ the skip copy into the decoder's concat buffer), the 64->2 head,
// the pixel-wise weighted losses, bias-gradient column sums, weight (un)packing and
// the fused SGD-momentum step.  All use 128-bit accesses on NHWC bf16 tensors:
// a thread owns 8 consecutive channels (16 B) of one pixel, so a warp touches
// 512 contiguous bytes.
#pragma once
#include "ptx.cuh"

namespace ub {

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// Ampere-style asynchronous copies (global -> shared without a register round trip): the staging mechanism of the
// two conv11c kernels, whose operands are too narrow / misaligned for TMA boxes (a 4-byte-aligned fp32 image).
__device__ __forceinline__ void cp_async_4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------
// conv11c: 1 -> 64 channels, 3x3 valid, fp32 image in, bf16 NHWC out, fused
// bias + ReLU (network.py:131).  K = 9 cannot feed the tensor pipe; the kernel
// is bound by its 2 B x 64 x pixels output write and by 9 fp32 FMAs per output.
// Thread = (pixel column, channel QUAD): 36 weights + 4 biases in registers, 8-byte stores, a warp writes the 256
// contiguous bytes of two pixels.  The block's image window ((rows + 2) x 18 floats) is staged in shared memory
// once: with global loads in the row loop the kernel sat in long-scoreboard stalls (ncu: 9.4 of 13.9 cycles per
// issued instruction) however many warps were resident.
constexpr int C1_ROWS = 96;   // output rows per block
__global__ void __launch_bounds__(256, 4) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                        int n_img, int H, int W) {
  // grid = (ceil(Wo/16), ceil(Ho/C1_ROWS), n): no index divisions; a thread keeps its channel quad and its x
  // position and walks down the rows; the three input rows slide (6 of 9 taps reused per step).
  __shared__ float sx[(C1_ROWS + 2) * 18];
  const int Ho = H - 2, Wo = W - 2;
  const int x0 = blockIdx.x * 16;
  const int px = threadIdx.x >> 4, q = threadIdx.x & 15;
  const int ox = x0 + px;
  const int n = blockIdx.z;
  const int y_begin = blockIdx.y * C1_ROWS;
  const int y_end = min(Ho, y_begin + C1_ROWS);
  const int rows_in = y_end - y_begin + 2;
  const float* xb = x + ((long long)n * H + y_begin) * W + x0;
  const int cols_in = min(18, W - x0);
  for (int i = threadIdx.x; i < rows_in * 18; i += 256) {
    const int r = i / 18, c = i - r * 18;
    if (c < cols_in) cp_async_4(&sx[i], xb + (long long)r * W + c);
  }
  cp_async_commit();
  // Packed fp32 pairs (fma.rn.f32x2, sm_100): one instruction updates two output channels, so the 36 FMAs per pixel
  // quad become 18 -- the scalar version was FMA-pipe bound (a three-register FFMA occupies its pipe for two cycles:
  // ncu 73 % issue-active, math-pipe throttle the top stall).  Same rounding as fmaf per element: bit-identical output.
  uint64_t wr2[2][9], br2[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    br2[c] = pack_f32x2(__ldg(b + q * 4 + 2 * c), __ldg(b + q * 4 + 2 * c + 1));
#pragma unroll
    for (int t = 0; t < 9; ++t) wr2[c][t] = pack_f32x2(__ldg(w + (q * 4 + 2 * c) * 9 + t), __ldg(w + (q * 4 + 2 * c + 1) * 9 + t));
  }
  cp_async_wait<0>();
  __syncthreads();
  if (ox >= Wo) return;
  const float* sp = sx + px;
  uint64_t in2[9];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    in2[3 + s] = pack_f32x2(sp[s], sp[s]);
    in2[6 + s] = pack_f32x2(sp[18 + s], sp[18 + s]);
  }
  __nv_bfloat16* yp = y + (((long long)n * Ho + y_begin) * Wo + ox) * 64 + q * 4;
  for (int r = 0; r < y_end - y_begin; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      in2[s] = in2[3 + s];
      in2[3 + s] = in2[6 + s];
      const float v = sp[(r + 2) * 18 + s];
      in2[6 + s] = pack_f32x2(v, v);
    }
    uint2 pk;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint64_t a = br2[c];
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fma_f32x2(in2[t], wr2[c][t], a);
      float lo, hi;
      unpack_f32x2(a, lo, hi);
      const uint32_t v = pack_bf16x2(fmaxf(lo, 0.0f), fmaxf(hi, 0.0f));
      if (c == 0) pk.x = v; else pk.y = v;
    }
    *reinterpret_cast<uint2*>(yp) = pk;
    yp += (long long)Wo * 64;
  }
}

// fp32 image [n][H][W] -> three bf16 copies [3][n*H][pitch], copy s shifted left by s pixels (copy_s[x] = image[x + s],
// zero beyond the row; pitch = W rounded up to 8 elements): the B operand source of the tensor-pipe conv11c weight
// gradient (wgrad.cuh, CONV1), whose TMA boxes need 16-byte aligned starts and row strides.
__global__ void __launch_bounds__(256) image_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            int W, int pitch, long long rows) {
  const long long plane = rows * pitch;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < plane; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / pitch;
    const int c = (int)(g - r * pitch);
    const float* xr = x + r * W;
#pragma unroll
    for (int s = 0; s < 3; ++s) y[s * plane + g] = __float2bfloat16_rn(c + s < W ? __ldg(xr + c + s) : 0.0f);
  }
}

// conv11c weight + bias gradient, CUDA-core version (kept as the comparator of the tensor-pipe one; ub_set_option(8, 0)):
// dW[co, tap] = sum_p dY[p, co] * x[p + tap], db[co] = sum_p dY[p, co].
// dY is bf16 NHWC (already ReLU-masked), 8 KB per 64-pixel tile, streamed through a 3-stage cp.async ring together
// with the tile's 3 x 66 image window, so ~16 KB per block (5 blocks per SM) are in flight whatever the compute
// does.  Thread = (pixel slot, channel octet, tap half): half 0 accumulates taps 0..4, half 1 taps 5..8 and the
// bias (x = 1): 40 accumulators.  Block-level reduction in shared memory (fixed order); every block STORES its 576 + 64
// partial sums as one row of `part` ([blocks][640]: dW[co][tap] then db[co]) -- no atomics; the rows are summed in
// order by the unpack kernel (deterministic gradients).
constexpr int C1W_TILE = 64, C1W_STAGES = 3;
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x,
                                                          const __nv_bfloat16* __restrict__ dy,
                                                          float* __restrict__ part, int n_img,
                                                          int H, int W, int rows_per_block) {
  __shared__ __align__(16) uint8_t sdy[C1W_STAGES][C1W_TILE * 128];
  __shared__ float sxw[C1W_STAGES][3][C1W_TILE + 2];
  __shared__ float sred[8][16][41];   // [warp][(half, octet)][channel-in-octet x 5 taps], padded
  const int Ho = H - 2, Wo = W - 2;
  const int n = blockIdx.y;
  const int oct = threadIdx.x & 7, half = (threadIdx.x >> 3) & 1, px = threadIdx.x >> 4;
  const int y_begin = blockIdx.x * rows_per_block;
  const int y_end = min(Ho, y_begin + rows_per_block);
  const int xtiles = (Wo + C1W_TILE - 1) / C1W_TILE;
  const int n_tiles = (y_end - y_begin) * xtiles;
  int soff[5];   // offsets of this half's taps inside the staged window; the 5th of half 1 is the bias pseudo-tap
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int t = min(half * 5 + j, 8);
    soff[j] = (t / 3) * (C1W_TILE + 2) + (t % 3);
  }
  float acc[8][5];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[c][j] = 0.0f;

  auto issue = [&](int t) {
    if (t < n_tiles) {
      const int st = t % C1W_STAGES;
      const int oy = y_begin + t / xtiles, x0 = (t % xtiles) * C1W_TILE;
      const int npx = min(C1W_TILE, Wo - x0);
      const __nv_bfloat16* g = dy + (((long long)n * Ho + oy) * Wo + x0) * 64;
      for (int i = threadIdx.x; i < npx * 8; i += 256) cp_async_16(&sdy[st][i * 16], g + i * 8);
      const float* xr = x + ((long long)n * H + oy) * W + x0;
      for (int i = threadIdx.x; i < 3 * (C1W_TILE + 2); i += 256) {
        const int r = i / (C1W_TILE + 2), c = i - r * (C1W_TILE + 2);
        if (c < npx + 2) cp_async_4(&sxw[st][r][c], xr + (long long)r * W + c);
      }
    }
    cp_async_commit();   // an (empty) group per call keeps the wait_group arithmetic uniform
  };
  for (int t = 0; t < C1W_STAGES - 1; ++t) issue(t);
  for (int t = 0; t < n_tiles; ++t) {
    cp_async_wait<C1W_STAGES - 2>();
    __syncthreads();                 // tile t has landed for everyone; everyone is done with tile t-1's stage
    issue(t + C1W_STAGES - 1);
    const int st = t % C1W_STAGES;
    const int npx = min(C1W_TILE, Wo - (t % xtiles) * C1W_TILE);
    const float* sw = &sxw[st][0][0];
#pragma unroll
    for (int i = 0; i < C1W_TILE / 16; ++i) {
      const int p = px + 16 * i;
      if (p < npx) {
        float g[8], in[5];
        unpack8(*reinterpret_cast<const uint4*>(&sdy[st][p * 128 + oct * 16]), g);
#pragma unroll
        for (int j = 0; j < 5; ++j) in[j] = sw[p + soff[j]];
        if (half) in[4] = 1.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int j = 0; j < 5; ++j) acc[c][j] = fmaf(g[c], in[j], acc[c][j]);
      }
    }
  }
  cp_async_wait<0>();
  // lanes with equal (lane & 15) share (octet, half): reduce over xor 16; then across warps through smem
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[c][j] += __shfl_xor_sync(0xffffffffu, acc[c][j], 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 5; ++j) sred[warp][half * 8 + oct][c * 5 + j] = acc[c][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 10; i += blockDim.x) {
    const int ch = i / 10, t = i % 10;
    float v = 0.0f;
    for (int wv = 0; wv < 8; ++wv) v += sred[wv][(t / 5) * 8 + (ch >> 3)][(ch & 7) * 5 + (t % 5)];
    float* row = part + (long long)(blockIdx.y * gridDim.x + blockIdx.x) * 640;
    if (t < 9)
      row[ch * 9 + t] = v;
    else
      row[576 + ch] = v;
  }
}

// ---------------------------------------------------------------------------
// 2x2 stride-2 max pool (network.py:133,139,145,151) on bf16 NHWC.
// Writes: dense pooled tensor, 1-byte argmax code (dy*2+dx, first maximum in
// row-major window order, NaN wins -- the ATen rule), and the same value into the
// decoder's concat buffer at the crop_and_concat offset (network.py:124-127):
// concat coordinate = pooled coordinate - c, c = int((Hp - Hu) * 0.5).
__global__ void __launch_bounds__(256) pool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                                                       __nv_bfloat16* __restrict__ cat, int n_img, int H, int W, int C,
                                                       int cat_H, int cat_W, int cat_C, int c_off) {
  const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
  const long long total = (long long)n_img * Ho * Wo * C8;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(gid % C8);
    long long pix = gid / C8;
    const int ox = (int)(pix % Wo);
    long long r2 = pix / Wo;
    const int oy = (int)(r2 % Ho);
    const int n = (int)(r2 / Ho);
    const __nv_bfloat16* xp = x + (((long long)n * H + 2 * oy) * W + 2 * ox) * C + c8 * 8;
    float v[4][8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xp)), v[0]);
    unpack8(__ldg(reinterpret_cast<const uint4*>(xp + C)), v[1]);
    unpack8(__ldg(reinterpret_cast<const uint4*>(xp + (long long)W * C)), v[2]);
    unpack8(__ldg(reinterpret_cast<const uint4*>(xp + (long long)W * C + C)), v[3]);
    float m[8];
    uint32_t code[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      m[c] = v[0][c];
      code[c] = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float t = v[k][c];
        if (t > m[c] || t != t) {
          m[c] = t;
          code[c] = k;
        }
      }
    }
    const uint4 o = pack8(m);
    *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) = o;
    uint2 cc;
    cc.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
    cc.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
    *reinterpret_cast<uint2*>(idx + pix * C + c8 * 8) = cc;
    if (cat != nullptr) {
      const int cy = oy - c_off, cx = ox - c_off;
      if (cy >= 0 && cy < cat_H && cx >= 0 && cx < cat_W)
        *reinterpret_cast<uint4*>(cat + (((long long)n * cat_H + cy) * cat_W + cx) * cat_C + c8 * 8) = o;
    }
  }
}

// Max-pool backward fused with (a) the sum of the two gradient producers of the pooled
// tensor -- the next encoder conv's dgrad (dense) and the skip half of the concat
// gradient (window of dcat at the crop_and_concat offset) -- and (b) the ReLU backward of
// the conv that fed the pool (gradient passes only where the pooled value > 0).
template <bool BIAS>
__global__ void __launch_bounds__(256, BIAS ? 6 : 7) pool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                       const __nv_bfloat16* __restrict__ dcat,
                                                       const uint8_t* __restrict__ idx,
                                                       const __nv_bfloat16* __restrict__ pooled,
                                                       __nv_bfloat16* __restrict__ dx, int n_img, int H, int W, int C,
                                                       int cat_H, int cat_W, int cat_C, int c_off,
                                                       float* __restrict__ dbias) {
  // BIAS: also accumulate the column sums of dx (= the bias gradient of the conv that fed the pool,
  // network.py:132-151) -- every pooled element contributes its one scattered value.  256 % C8 == 0 (checked by
  // the host), so a thread keeps its channel octet over the whole grid-stride loop and sums in registers.
  const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
  const long long total = (long long)n_img * Ho * Wo * C8;
  float bsum[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bsum[c] = 0.0f;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(gid % C8);
    long long pix = gid / C8;
    const int ox = (int)(pix % Wo);
    long long r2 = pix / Wo;
    const int oy = (int)(r2 % Ho);
    const int n = (int)(r2 / Ho);
    float g[8], pv[8];
    if (dy != nullptr) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + pix * C + c8 * 8)), g);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) g[c] = 0.0f;
    }
    if (dcat != nullptr) {
      const int cy = oy - c_off, cx = ox - c_off;
      if (cy >= 0 && cy < cat_H && cx >= 0 && cx < cat_W) {
        float g2[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dcat + (((long long)n * cat_H + cy) * cat_W + cx) * cat_C + c8 * 8)),
                g2);
#pragma unroll
        for (int c = 0; c < 8; ++c) g[c] += g2[c];
      }
    }
    unpack8(__ldg(reinterpret_cast<const uint4*>(pooled + pix * C + c8 * 8)), pv);
    const uint2 cc = __ldg(reinterpret_cast<const uint2*>(idx + pix * C + c8 * 8));
    float o[4][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t code = ((c < 4 ? cc.x : cc.y) >> ((c & 3) * 8)) & 0xff;
      const float gv = pv[c] > 0.0f ? g[c] : 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][c] = (code == (uint32_t)k) ? gv : 0.0f;
      if (BIAS) bsum[c] += gv;
    }
    __nv_bfloat16* xp = dx + (((long long)n * H + 2 * oy) * W + 2 * ox) * C + c8 * 8;
    *reinterpret_cast<uint4*>(xp) = pack8(o[0]);
    *reinterpret_cast<uint4*>(xp + C) = pack8(o[1]);
    *reinterpret_cast<uint4*>(xp + (long long)W * C) = pack8(o[2]);
    *reinterpret_cast<uint4*>(xp + (long long)W * C + C) = pack8(o[3]);
  }
  if (BIAS) {
    __shared__ float sred[256][9];
#pragma unroll
    for (int c = 0; c < 8; ++c) sred[threadIdx.x][c] = bsum[c];
    __syncthreads();
    const int slots = 256 / C8;   // threads tid, tid + C8, ... share the octet tid % C8
    for (int i = threadIdx.x; i < C; i += 256) {
      const int cc8 = i >> 3, c = i & 7;
      float v = 0.0f;
      for (int sl = 0; sl < slots; ++sl) v += sred[sl * C8 + cc8][c];
      dbias[(long long)blockIdx.x * C + i] = v;   // row blockIdx.x of the partial sums [gridDim.x][C] (no atomics)
    }
  }
}

// ---------------------------------------------------------------------------
// finalconv: 1x1, 64 -> 2 (network.py:190).  bf16 NHWC in, fp32 NCHW logits out,
// optional fused 2-class argmax (tester.py:30, trainer.py:82: ties -> class 0).
// 8 threads per pixel, each 8 channels, shuffle-reduced.
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ logits, uint8_t* __restrict__ amax,
                                                       long long n_pix_per_img, int n_img) {
  const int oct = threadIdx.x & 7;
  float w0[8], w1[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    w0[c] = __ldg(w + oct * 8 + c);
    w1[c] = __ldg(w + 64 + oct * 8 + c);
  }
  const float b0 = __ldg(b), b1 = __ldg(b + 1);
  const unsigned int total = (unsigned int)(n_pix_per_img * n_img * 8);
  const unsigned int total_r = (total + 31u) & ~31u;  // keep whole warps in the loop for the shuffles
  const unsigned int npp = (unsigned int)n_pix_per_img;
  for (unsigned int gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total_r; gid += gridDim.x * blockDim.x) {
    const unsigned int pix = gid >> 3;
    float s0 = 0.0f, s1 = 0.0f;
    if (gid < total) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (long long)pix * 64 + oct * 8)), v);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s0 = fmaf(v[c], w0[c], s0);
        s1 = fmaf(v[c], w1[c], s1);
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (gid < total && oct == 0) {
      s0 += b0;
      s1 += b1;
      const unsigned int n = pix / npp, r = pix - n * npp;
      if (logits != nullptr) {
        logits[(long long)(n * 2) * npp + r] = s0;
        logits[(long long)(n * 2 + 1) * npp + r] = s1;
      }
      if (amax != nullptr) amax[pix] = (s1 > s0) ? 1 : 0;
    }
  }
}

// Head backward: dX[p, c] = (dl0[p] W0[c] + dl1[p] W1[c]) * (X[p, c] > 0)  (ReLU of conv12e),
// dW[k, c] = sum_p dl_k[p] X[p, c], db[k] = sum_p dl_k[p].
__global__ void __launch_bounds__(256) head_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const float* __restrict__ w,
                                                       const float* __restrict__ dlogits,
                                                       __nv_bfloat16* __restrict__ dx, float* __restrict__ part,
                                                       int with_dx_sums, long long n_pix_per_img, int n_img) {
  // part: partial sums [gridDim.x][194] = dW[2][64], db[2] and (with_dx_sums) the column sums of dX = the bias
  // gradient of conv12e (network.py:188); summed in row order by the unpack kernel
  const int oct = threadIdx.x & 7;
  float w0[8], w1[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    w0[c] = __ldg(w + oct * 8 + c);
    w1[c] = __ldg(w + 64 + oct * 8 + c);
  }
  float a0[8], a1[8], bs[8], sb0 = 0.0f, sb1 = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) a0[c] = a1[c] = bs[c] = 0.0f;
  const unsigned int total = (unsigned int)(n_pix_per_img * n_img * 8);
  const unsigned int npp = (unsigned int)n_pix_per_img;
  for (unsigned int gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += gridDim.x * blockDim.x) {
    const unsigned int pix = gid >> 3;
    const unsigned int n = pix / npp, r = pix - n * npp;
    const float d0 = __ldg(dlogits + (long long)(n * 2) * npp + r);
    const float d1 = __ldg(dlogits + (long long)(n * 2 + 1) * npp + r);
    float v[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (long long)pix * 64 + oct * 8)), v);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      o[c] = v[c] > 0.0f ? fmaf(d0, w0[c], d1 * w1[c]) : 0.0f;
      bs[c] += __bfloat162float(__float2bfloat16_rn(o[c]));
      a0[c] = fmaf(d0, v[c], a0[c]);
      a1[c] = fmaf(d1, v[c], a1[c]);
    }
    if (oct == 0) {
      sb0 += d0;
      sb1 += d1;
    }
    *reinterpret_cast<uint4*>(dx + (long long)pix * 64 + oct * 8) = pack8(o);
  }
  // reduce lanes with equal oct (xor 8, 16), then across warps through smem
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    a0[c] += __shfl_xor_sync(0xffffffffu, a0[c], 8);
    a0[c] += __shfl_xor_sync(0xffffffffu, a0[c], 16);
    a1[c] += __shfl_xor_sync(0xffffffffu, a1[c], 8);
    a1[c] += __shfl_xor_sync(0xffffffffu, a1[c], 16);
    bs[c] += __shfl_xor_sync(0xffffffffu, bs[c], 8);
    bs[c] += __shfl_xor_sync(0xffffffffu, bs[c], 16);
  }
  sb0 = warp_sum(sb0);
  sb1 = warp_sum(sb1);
  __shared__ float sred[8][194];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      sred[warp][lane * 8 + c] = a0[c];
      sred[warp][64 + lane * 8 + c] = a1[c];
      sred[warp][130 + lane * 8 + c] = bs[c];
    }
  }
  if (lane == 0) {
    sred[warp][128] = sb0;
    sred[warp][129] = sb1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (with_dx_sums ? 194 : 130); i += blockDim.x) {
    float v = 0.0f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sred[wv][i];
    part[(long long)blockIdx.x * 194 + i] = v;   // row of partial sums: dW[2][64], db[2], column sums of dx [64]
  }
}

// ---------------------------------------------------------------------------
// Pixel-wise weighted losses on 2-channel fp32 logits [B,2,Hp,Wp] (centre-cropped
// to the label size: trainer.py:60-61), label [B,L,L] (uint8 or int64), weight
// [B,L,L] fp32.  kind 0: BCE-with-logits on the one-hot target, mean over B*2*L*L
// (trainer.py:63-75, nn.BCEWithLogitsLoss(weight)); kind 1: weighted softmax
// cross-entropy, mean over B*L*L (BASELINE config 5).  Fused forward + dlogits.
// Deterministic reduction: per-block partials, last block sums them in order.
template <typename LabelT>
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ logits, const LabelT* __restrict__ label,
                                                   const float* __restrict__ weight, float* __restrict__ dlogits,
                                                   float* __restrict__ loss_out, float* __restrict__ partials,
                                                   unsigned int* __restrict__ counter, int B, int Hp, int Wp, int L,
                                                   int pad, int kind, float grad_scale, int vec4) {
  // A thread owns 4 consecutive pixels of one row (vec4 = 1: 128-bit loads/stores; vec4 = 0: the same
  // loop with scalar accesses when L, pad or the row pitch break 16-byte alignment).
  const unsigned int qpr = (unsigned int)(L + 3) >> 2;            // quads per row
  const unsigned int total_q = (unsigned int)B * L * qpr;
  const long long plane = (long long)Hp * Wp;
  float acc = 0.0f;
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < total_q; q += gridDim.x * blockDim.x) {
    const unsigned int xq = q % qpr;
    const unsigned int r = q / qpr;
    const unsigned int yy = r % (unsigned int)L;
    const unsigned int n = r / (unsigned int)L;
    const int xx = (int)xq * 4;
    const long long off = (long long)(yy + pad) * Wp + (xx + pad);
    const long long lin = ((long long)n * L + yy) * L + xx;
    const float* z0p = logits + (n * 2ll) * plane + off;
    const float* z1p = z0p + plane;
    float z0[4], z1[4], wv[4], g0[4], g1[4];
    int lab[4];
    const int cnt = min(4, L - xx);
    if (vec4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(z0p));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(z1p));
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(weight + lin));
      z0[0] = a.x; z0[1] = a.y; z0[2] = a.z; z0[3] = a.w;
      z1[0] = b4.x; z1[1] = b4.y; z1[2] = b4.z; z1[3] = b4.w;
      wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
#pragma unroll
      for (int i = 0; i < 4; ++i) lab[i] = (int)label[lin + i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = i < cnt;
        z0[i] = ok ? __ldg(z0p + i) : 0.0f;
        z1[i] = ok ? __ldg(z1p + i) : 0.0f;
        wv[i] = ok ? __ldg(weight + lin + i) : 0.0f;
        lab[i] = ok ? (int)label[lin + i] : 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (kind == 0) {
        // softplus(z) - y z = max(z,0) - y z + log1p(exp(-|z|));  sigmoid from the same exponential
        const float y0 = lab[i] ? 0.0f : 1.0f, y1 = lab[i] ? 1.0f : 0.0f;
        const float e0 = __expf(-fabsf(z0[i])), e1 = __expf(-fabsf(z1[i]));
        // e in (0, 1]: log(1 + e) through the fast log keeps ~1e-7 absolute error on a term that is >= e/2
        const float l0 = fmaxf(z0[i], 0.0f) - y0 * z0[i] + __logf(1.0f + e0);
        const float l1 = fmaxf(z1[i], 0.0f) - y1 * z1[i] + __logf(1.0f + e1);
        acc += wv[i] * (l0 + l1);
        const float r0 = __fdividef(1.0f, 1.0f + e0), r1 = __fdividef(1.0f, 1.0f + e1);
        const float s0 = z0[i] >= 0.0f ? r0 : e0 * r0, s1 = z1[i] >= 0.0f ? r1 : e1 * r1;
        g0[i] = wv[i] * (s0 - y0) * grad_scale;
        g1[i] = wv[i] * (s1 - y1) * grad_scale;
      } else {
        const float m = fmaxf(z0[i], z1[i]);
        const float e0 = __expf(z0[i] - m), e1 = __expf(z1[i] - m);
        const float ssum = e0 + e1;   // in [1, 2]
        acc += wv[i] * (m + __logf(ssum) - (lab[i] ? z1[i] : z0[i]));
        const float inv = __fdividef(1.0f, ssum);
        g0[i] = wv[i] * (e0 * inv - (lab[i] ? 0.0f : 1.0f)) * grad_scale;
        g1[i] = wv[i] * (e1 * inv - (lab[i] ? 1.0f : 0.0f)) * grad_scale;
      }
    }
    if (dlogits != nullptr) {
      // dlogits has the logits' own (uncropped) shape; the cropped-away border stays zero
      float* d0p = dlogits + (n * 2ll) * plane + off;
      float* d1p = d0p + plane;
      if (vec4) {
        *reinterpret_cast<float4*>(d0p) = make_float4(g0[0], g0[1], g0[2], g0[3]);
        *reinterpret_cast<float4*>(d1p) = make_float4(g1[0], g1[1], g1[2], g1[3]);
      } else {
        for (int i = 0; i < cnt; ++i) {
          d0p[i] = g0[i];
          d1p[i] = g1[i];
        }
      }
    }
  }
  acc = warp_sum(acc);
  __shared__ float sw[8];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sw[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += sw[i];
    partials[blockIdx.x] = v;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order tree over the partials: deterministic for a given grid size
    float v = 0.0f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += partials[i];
    v = warp_sum(v);
    if (lane == 0) sw[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sw[i];
      const float npx = (float)B * (float)L * (float)L;
      const float denom = (kind == 0) ? 2.0f * npx : npx;
      *loss_out = t / denom;
      *counter = 0;  // re-arm for the next launch (graph replays included)
    }
  }
}

// Fast path of the same losses for the common geometry -- no crop (pad == 0, logits plane == label plane, as
// in the network's own 572 -> 388 step and BASELINE config 5) and a plane of a multiple of 8 pixels: each image
// plane is one linear run, so a thread owns 8 consecutive pixels and issues all of its loads (2 x 128 bit per
// logit channel and weight, 64 bit of uint8 labels) before the first use: ~104 B in flight per thread instead
// of 52, no row/column index arithmetic.  Same partial-sum protocol as loss_kernel.
// 8 labels -> one bit mask (bit i = label i is non-zero, i.e. class 1 of the one-hot target, trainer.py:63-66)
__device__ __forceinline__ uint32_t load_labels8(const uint8_t* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((u.x >> (8 * i)) & 0xffu) ? (1u << i) : 0u;
    m |= ((u.y >> (8 * i)) & 0xffu) ? (16u << i) : 0u;
  }
  return m;
}
__device__ __forceinline__ uint32_t load_labels8(const long long* p) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(p) + i);
    m |= (v.x != 0 ? 1u : 0u) << (2 * i);
    m |= (v.y != 0 ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}

template <typename LabelT>
struct LossUnit {           // the 8 pixels a thread works on: 6 x 128-bit loads + the labels
  float4 a0, a1, b0, b1, w0, w1;
  uint32_t lab;   // bit i = class of pixel i
};

template <typename LabelT>
__device__ __forceinline__ void loss_unit_load(LossUnit<LabelT>& q, unsigned int u, unsigned int plane8,
                                               long long plane, const float* __restrict__ logits,
                                               const LabelT* __restrict__ label, const float* __restrict__ weight) {
  const unsigned int n = u / plane8;
  const long long loff = (long long)u * 8;
  const long long zoff = loff + (long long)n * plane;    // = n*2*plane + (u - n*plane8)*8
  const float4* z0p = reinterpret_cast<const float4*>(logits + zoff);
  const float4* z1p = reinterpret_cast<const float4*>(logits + zoff + plane);
  const float4* wp = reinterpret_cast<const float4*>(weight + loff);
  q.a0 = __ldg(z0p); q.a1 = __ldg(z0p + 1);
  q.b0 = __ldg(z1p); q.b1 = __ldg(z1p + 1);
  q.w0 = __ldg(wp);  q.w1 = __ldg(wp + 1);
  q.lab = load_labels8(label + loff);
}

// Persistent grid (4 CTAs per SM), grid-stride over 8-pixel units with a one-unit register prefetch: the loads
// of unit i+1 are in flight while the ~50 MUFU operations of unit i run, so the SFU time (6 MUFU per pixel for
// the BCE form, ~12 us at config 5) hides under the HBM time instead of adding to it.
template <typename LabelT, bool WANT_GRAD>
__global__ void __launch_bounds__(256, WANT_GRAD ? 3 : 4) loss_linear_kernel(const float* __restrict__ logits,
                                                             const LabelT* __restrict__ label,
                                                             const float* __restrict__ weight,
                                                             float* __restrict__ dlogits, float* __restrict__ loss_out,
                                                             float* __restrict__ partials,
                                                             unsigned int* __restrict__ counter, int B,
                                                             unsigned int plane8, int kind, float grad_scale) {
  const unsigned int total_u = (unsigned int)B * plane8;
  const long long plane = (long long)plane8 * 8;
  const unsigned int stride = gridDim.x * blockDim.x;
  constexpr bool want_grad = WANT_GRAD;
  float acc = 0.0f;
  unsigned int u = blockIdx.x * blockDim.x + threadIdx.x;
  LossUnit<LabelT> nx;
  if (u < total_u) loss_unit_load(nx, u, plane8, plane, logits, label, weight);
  for (; u < total_u; u += stride) {
    const LossUnit<LabelT> q = nx;
    if (u + stride < total_u) loss_unit_load(nx, u + stride, plane8, plane, logits, label, weight);
    const float z0[8] = {q.a0.x, q.a0.y, q.a0.z, q.a0.w, q.a1.x, q.a1.y, q.a1.z, q.a1.w};
    const float z1[8] = {q.b0.x, q.b0.y, q.b0.z, q.b0.w, q.b1.x, q.b1.y, q.b1.z, q.b1.w};
    const float wv[8] = {q.w0.x, q.w0.y, q.w0.z, q.w0.w, q.w1.x, q.w1.y, q.w1.z, q.w1.w};
    float g0[8], g1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool lab = (q.lab >> i) & 1u;
      if (kind == 0) {
        const float y0 = lab ? 0.0f : 1.0f, y1 = lab ? 1.0f : 0.0f;
        const float e0 = __expf(-fabsf(z0[i])), e1 = __expf(-fabsf(z1[i]));
        const float l0 = fmaxf(z0[i], 0.0f) - y0 * z0[i] + __logf(1.0f + e0);
        const float l1 = fmaxf(z1[i], 0.0f) - y1 * z1[i] + __logf(1.0f + e1);
        acc += wv[i] * (l0 + l1);
        if (want_grad) {   // forward-only calls (validation) skip the two reciprocals: 4 instead of 6 MUFU per pixel
          const float r0 = __fdividef(1.0f, 1.0f + e0), r1 = __fdividef(1.0f, 1.0f + e1);
          const float s0 = z0[i] >= 0.0f ? r0 : e0 * r0, s1 = z1[i] >= 0.0f ? r1 : e1 * r1;
          g0[i] = wv[i] * (s0 - y0) * grad_scale;
          g1[i] = wv[i] * (s1 - y1) * grad_scale;
        }
      } else {
        // two classes: logsumexp(z) - z_l = softplus(z_other - z_l), softmax_1 = sigmoid(z1 - z0): 3 MUFU per pixel
        const float d = z1[i] - z0[i];
        const float e = __expf(-fabsf(d));
        const float r = __fdividef(1.0f, 1.0f + e);
        const float p1 = d >= 0.0f ? r : e * r;               // softmax probability of class 1
        const float t = lab ? -d : d;                          // z_other - z_label
        acc += wv[i] * (fmaxf(t, 0.0f) + __logf(1.0f + e));
        if (want_grad) {
          g1[i] = wv[i] * (p1 - (lab ? 1.0f : 0.0f)) * grad_scale;
          g0[i] = -g1[i];
        }
      }
    }
    if (WANT_GRAD) {
      const long long zoff = (long long)u * 8 + (long long)(u / plane8) * plane;
      float4* d0p = reinterpret_cast<float4*>(dlogits + zoff);
      float4* d1p = reinterpret_cast<float4*>(dlogits + zoff + plane);
      d0p[0] = make_float4(g0[0], g0[1], g0[2], g0[3]);
      d0p[1] = make_float4(g0[4], g0[5], g0[6], g0[7]);
      d1p[0] = make_float4(g1[0], g1[1], g1[2], g1[3]);
      d1p[1] = make_float4(g1[4], g1[5], g1[6], g1[7]);
    }
  }
  acc = warp_sum(acc);
  __shared__ float sw[8];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sw[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += sw[i];
    partials[blockIdx.x] = v;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float v = 0.0f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += partials[i];
    v = warp_sum(v);
    if (lane == 0) sw[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sw[i];
      const float npx = (float)B * (float)plane;
      *loss_out = t / ((kind == 0) ? 2.0f * npx : npx);
      *counter = 0;
    }
  }
}

// ---------------------------------------------------------------------------
// Column sums of a bf16 [pixels, C] view (pixel stride `ld` elements): bias gradients.
// blockDim = 256 = 32 pixel lanes x 8 ... generalised: thread = (pixel slot, channel octet).
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ out,
                                                     long long n_pix, int C, long long ld) {
  const int C8 = C >> 3;
  // threads are laid out [pixel_slot][c8] with c8 fastest; requires 256 % C8 == 0 or C8 % 256 == 0
  const int tpb = blockDim.x;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
  if (C8 <= tpb) {
    const int slots = tpb / C8;
    const int c8 = threadIdx.x % C8, slot = threadIdx.x / C8;
    for (long long pix = (long long)blockIdx.x * slots + slot; pix < n_pix; pix += (long long)gridDim.x * slots) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(g + pix * ld + c8 * 8)), v);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += v[c];
    }
    __shared__ float sred[256][9];
#pragma unroll
    for (int c = 0; c < 8; ++c) sred[threadIdx.x][c] = acc[c];
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += tpb) {
      const int cc8 = i >> 3, c = i & 7;
      float v = 0.0f;
      for (int s = 0; s < slots; ++s) v += sred[s * C8 + cc8][c];
      out[(long long)blockIdx.x * C + i] = v;     // row blockIdx.x of the partial sums [gridDim.x][C] (no atomics)
    }
  } else {
    // C8 > 256 never happens for this network (C <= 1024 -> C8 <= 128)
  }
}

// ---------------------------------------------------------------------------
// Weight packing: fp32 masters in torch layout -> bf16 K-major GEMM operands.
//  mode 0 conv fwd   : dst[co][(r,s)][ci]            = W[co][ci][r][s]
//  mode 1 conv dgrad : dst[ci][(r',s')][co]          = W[co][ci][2-r'][2-s']
//  mode 2 upconv fwd : dst[(a,b,co)][ci]             = W[ci][co][a][b]
//  mode 3 upconv dgrad: dst[ci][(a,b,co)]            = W[ci][co][a][b]
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst,
                                                          int mode, int Cout, int Cin) {
  const int T = (mode < 2) ? 9 : 4;
  const long long total = (long long)Cout * Cin * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long src;
    if (mode == 0) {
      const int ci = (int)(i % Cin);
      long long r = i / Cin;
      const int t = (int)(r % 9);
      const int co = (int)(r / 9);
      src = ((long long)co * Cin + ci) * 9 + t;
    } else if (mode == 1) {
      const int co = (int)(i % Cout);
      long long r = i / Cout;
      const int t = (int)(r % 9);
      const int ci = (int)(r / 9);
      src = ((long long)co * Cin + ci) * 9 + (8 - t);
    } else if (mode == 2) {
      const int ci = (int)(i % Cin);
      long long r = i / Cin;
      const int co = (int)(r % Cout);
      const int ab = (int)(r / Cout);
      src = ((long long)ci * Cout + co) * 4 + ab;
    } else {
      const int co = (int)(i % Cout);
      long long r = i / Cout;
      const int ab = (int)(r % 4);
      const int ci = (int)(r / 4);
      src = ((long long)ci * Cout + co) * 4 + ab;
    }
    dst[i] = __float2bfloat16_rn(__ldg(w + src));
  }
}

// All layers in ONE launch: a device-side job table replaces 42 (pack) / 21 (unpack) tiny launches per step.
struct PackJob {
  const float* w;          // fp32 master / gradient in torch layout
  __nv_bfloat16* dst_a;    // forward operand   (pack)
  __nv_bfloat16* dst_b;    // dgrad operand     (pack)
  const float* packed;     // packed fp32 gradient (unpack source)
  float* dw;               // torch-layout gradient (unpack destination)
  int kind;                // 0 = 3x3 conv, 2 = 2x2 up-conv, 3 = vector of `cout` elements (partial rows -> bias gradient)
  int cout, cin;
  long long elem_begin;    // prefix sum of element counts over the jobs
  int tile_begin;          // prefix sum of pack tile counts (pack_all)
  int ptile;               // pack tile edge in channels: 64 (both channel counts multiples of 64) or 32
  int slabs;               // split-K slabs the wgrad kernel of this layer leaves (unpack source = their sum)
  int lanes;               // 1, 2, 4 or 8 slab lanes per unpack tile (CO_T = 8 / lanes channels per tile)
  int wide;                // 1: conv job with < 8 slabs and cin % 128 == 0 -- unpack tile = 8 channels x 128 input channels
  int utile_begin;         // prefix sum of unpack tile counts: (cout / CO_T) * (cin / 32)
  long long slab_stride;   // elements between slabs
};

// Tiled through shared memory so that the fp32 side is read/written in long contiguous runs and the
// packed side in 64-128 B segments (the naive gather moved 8x the bytes through 32 B sectors).
//   conv   tile = 32 co x 32 ci x 9 taps   (W[co][ci][t]: 288 contiguous floats per co)
//   upconv tile = 32 ci x 32 co x 4 taps   (W[ci][co][ab]: 128 contiguous floats per ci)
__device__ __forceinline__ int find_tile_job(const PackJob* jobs, int n_jobs, int tile) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// 64 x 64-channel tiles staged as bf16 (pitch 578 / 258 elements): both packed operands leave in 128 B runs -- with
// 32 x 32 fp32 tiles the runs were 64 B and the kernel reached a third of the copy bandwidth.  Layers whose channel
// counts are not multiples of 64 (tests) keep the 32 x 32 path.
constexpr int PACK_SMEM_BYTES = 64 * 578 * 2;
__global__ void __launch_bounds__(256) pack_all_kernel(const PackJob* __restrict__ jobs, int n_jobs, int n_tiles) {
  extern __shared__ __align__(16) uint8_t pack_smem[];
  float* T = reinterpret_cast<float*>(pack_smem);
  __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(pack_smem);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const PackJob jb = jobs[find_tile_job(jobs, n_jobs, tile)];
    const int lt = tile - jb.tile_begin;
    const int Cin = jb.cin, Cout = jb.cout;
    __syncthreads();
    if (jb.ptile == 64 && jb.kind == 0) {
      const int ci_tiles = Cin >> 6;
      const int co0 = (lt / ci_tiles) * 64, ci0 = (lt % ci_tiles) * 64;
      for (int e = threadIdx.x; e < 64 * 144; e += 256) {      // 576 contiguous floats per co, float4 loads
        const int co = e / 144, k4 = e - co * 144;
        const float4 v = __ldg(reinterpret_cast<const float4*>(jb.w + ((long long)(co0 + co) * Cin + ci0) * 9) + k4);
        __nv_bfloat16* h = H + co * 578 + k4 * 4;
        *reinterpret_cast<__nv_bfloat162*>(h) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(h + 2) = __floats2bfloat162_rn(v.z, v.w);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 64 * 9 * 32; e += 256) {   // forward operand [co][t][ci]: a warp writes 128 B
        const int ci = (e & 31) * 2, r = e >> 5;
        const int t = r % 9, co = r / 9;
        __nv_bfloat162 v;
        v.x = H[co * 578 + ci * 9 + t];
        v.y = H[co * 578 + (ci + 1) * 9 + t];
        *reinterpret_cast<__nv_bfloat162*>(jb.dst_a + ((long long)(co0 + co) * 9 + t) * Cin + ci0 + ci) = v;
      }
      for (int e = threadIdx.x; e < 64 * 9 * 32; e += 256) {   // dgrad operand [ci][8-t][co]
        const int co = (e & 31) * 2, r = e >> 5;
        const int t = r % 9, ci = r / 9;
        __nv_bfloat162 v;
        v.x = H[co * 578 + ci * 9 + t];
        v.y = H[(co + 1) * 578 + ci * 9 + t];
        *reinterpret_cast<__nv_bfloat162*>(jb.dst_b + ((long long)(ci0 + ci) * 9 + (8 - t)) * Cout + co0 + co) = v;
      }
    } else if (jb.ptile == 64) {
      const int co_tiles = Cout >> 6;
      const int ci0 = (lt / co_tiles) * 64, co0 = (lt % co_tiles) * 64;
      for (int e = threadIdx.x; e < 64 * 64; e += 256) {       // 256 contiguous floats per ci
        const int ci = e >> 6, k4 = e & 63;
        const float4 v = __ldg(reinterpret_cast<const float4*>(jb.w + ((long long)(ci0 + ci) * Cout + co0) * 4) + k4);
        __nv_bfloat16* h = H + ci * 258 + k4 * 4;
        *reinterpret_cast<__nv_bfloat162*>(h) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(h + 2) = __floats2bfloat162_rn(v.z, v.w);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 4 * 64 * 32; e += 256) {   // forward operand [(ab,co)][ci]
        const int ci = (e & 31) * 2, r = e >> 5;
        const int co = r & 63, ab = r >> 6;
        __nv_bfloat162 v;
        v.x = H[ci * 258 + co * 4 + ab];
        v.y = H[(ci + 1) * 258 + co * 4 + ab];
        *reinterpret_cast<__nv_bfloat162*>(jb.dst_a + ((long long)ab * Cout + co0 + co) * Cin + ci0 + ci) = v;
      }
      for (int e = threadIdx.x; e < 64 * 4 * 32; e += 256) {   // dgrad operand [ci][(ab,co)]
        const int co = (e & 31) * 2, r = e >> 5;
        const int ab = r & 3, ci = r >> 2;
        __nv_bfloat162 v;
        v.x = H[ci * 258 + co * 4 + ab];
        v.y = H[ci * 258 + (co + 1) * 4 + ab];
        *reinterpret_cast<__nv_bfloat162*>(jb.dst_b + ((long long)(ci0 + ci) * 4 + ab) * Cout + co0 + co) = v;
      }
    } else if (jb.kind == 0) {
      const int ci_tiles = Cin >> 5;
      const int co0 = (lt / ci_tiles) * 32, ci0 = (lt % ci_tiles) * 32;
      for (int e = threadIdx.x; e < 32 * 288; e += 256) {
        const int co = e / 288, k = e - co * 288;
        T[co * 289 + k] = __ldg(jb.w + ((long long)(co0 + co) * Cin + ci0) * 9 + k);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * 288; e += 256) {      // forward operand [co][t][ci]
        const int ci = e & 31, r = e >> 5;
        const int t = r % 9, co = r / 9;
        jb.dst_a[((long long)(co0 + co) * 9 + t) * Cin + ci0 + ci] = __float2bfloat16_rn(T[co * 289 + ci * 9 + t]);
      }
      for (int e = threadIdx.x; e < 32 * 288; e += 256) {      // dgrad operand [ci][8-t][co]
        const int co = e & 31, r = e >> 5;
        const int t = r % 9, ci = r / 9;
        jb.dst_b[((long long)(ci0 + ci) * 9 + (8 - t)) * Cout + co0 + co] = __float2bfloat16_rn(T[co * 289 + ci * 9 + t]);
      }
    } else {
      const int co_tiles = Cout >> 5;
      const int ci0 = (lt / co_tiles) * 32, co0 = (lt % co_tiles) * 32;
      for (int e = threadIdx.x; e < 32 * 128; e += 256) {
        const int ci = e >> 7, k = e & 127;
        T[ci * 129 + k] = __ldg(jb.w + ((long long)(ci0 + ci) * Cout + co0) * 4 + k);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * 128; e += 256) {      // forward operand [(ab,co)][ci]
        const int ci = e & 31, r = e >> 5;
        const int co = r & 31, ab = r >> 5;
        jb.dst_a[((long long)ab * Cout + co0 + co) * Cin + ci0 + ci] = __float2bfloat16_rn(T[ci * 129 + co * 4 + ab]);
      }
      for (int e = threadIdx.x; e < 32 * 128; e += 256) {      // dgrad operand [ci][(ab,co)]
        const int co = e & 31, r = e >> 5;
        const int ab = r & 3, ci = r >> 2;
        jb.dst_b[((long long)(ci0 + ci) * 4 + ab) * Cout + co0 + co] = __float2bfloat16_rn(T[ci * 129 + co * 4 + ab]);
      }
    }
  }
}

// Gradient slabs -> torch layout, all layers (or a range of them) in one launch.  The wgrad kernels leave `slabs`
// partial gradients per layer (deterministic split-K: split s stored slab s, `slab_stride` elements apart); this
// kernel sums them IN FIXED ORDER and transposes packed -> torch layout through shared memory:
//   conv    packed [co][t][ci]    -> dW[co][ci][t]    tile = CO_T co x 32 ci x 9 taps  (288 contiguous floats per co)
//   up-conv packed [(ab,co)][ci]  -> dW[ci][co][ab]   tile = 32 ci x CO_T co x 4 taps
// 256 threads = 32 ci x 8 rows; the 8 rows are CO_T channels x LANES slab lanes (CO_T * LANES = 8): lane l sums
// slabs l, l + LANES, ... and the lanes are added in order, so layers with hundreds of slabs (the wide shallow
// ones, whose K dimension is split across the whole chip) still expose enough parallel loads.
constexpr int UNPACK_MAX_JOBS = 128;
constexpr int UP_PITCH = 33;   // odd pitch: the transposing read of the store phase is bank-conflict free
constexpr int UW_PITCH = 132;  // wide path: 128 input channels per row, 16-byte aligned rows
__global__ void __launch_bounds__(256, 2) unpack_all_kernel(const PackJob* __restrict__ jobs, int n_jobs, int tile_lo,
                                                         int tile_hi, int accumulate) {
  __shared__ __align__(16) float P[8 * 9 * UW_PITCH];   // block path: [8 rows][9 taps][UP_PITCH]; wide path: per warp [9][UW_PITCH]
  __shared__ int sbegin[UNPACK_MAX_JOBS];
  for (int i = threadIdx.x; i < n_jobs && i < UNPACK_MAX_JOBS; i += 256) sbegin[i] = jobs[i].utile_begin;
  __syncthreads();
  const int ci = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int n_tiles = tile_hi - tile_lo;
  bool prev_shared = false;   // the previous tile was read across warps: a block barrier before P is written again
  // Tiles of the first and of the last layers of the table (the wide shallow ones: up to 148 slabs to add per tile) are
  // the long ones; walking the range from both ends towards the middle starts them first, and the short single-slab
  // tiles of the deep layers fill the tail.
  for (int k = blockIdx.x; k < n_tiles; k += gridDim.x) {
    const int tile = (k & 1) ? tile_hi - 1 - (k >> 1) : tile_lo + (k >> 1);
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sbegin[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    const PackJob jb = jobs[lo];
    const int lt = tile - jb.utile_begin;
    const int Cin = jb.cin, Cout = jb.cout;
    if (jb.kind == 3) {
      // vector job (bias gradients, finalconv / conv11c weights): dw[e] = sum over the `slabs` partial rows the
      // producing kernel's CTAs stored; 32 elements per tile, 8 row lanes added in order
      prev_shared = true;
      __syncthreads();
      const int e = lt * 32 + ci;
      float v = 0.0f;
      if (e < Cout) {
#pragma unroll 4
        for (int s = y; s < jb.slabs; s += 8) v += __ldcg(jb.packed + (long long)s * jb.slab_stride + e);
      }
      P[y * UP_PITCH + ci] = v;
      __syncthreads();
      if (y == 0 && e < Cout) {
        float t = P[ci];
#pragma unroll
        for (int l = 1; l < 8; ++l) t += P[l * UP_PITCH + ci];
        jb.dw[e] = accumulate ? jb.dw[e] + t : t;
      }
      continue;
    }
    const int lanes = jb.lanes, co_t = 8 / lanes;
    const int ci_tiles = Cin >> 5;
    const int co0 = (lt / ci_tiles) * co_t, ci0 = (lt % ci_tiles) * 32;
    const int T = jb.kind == 0 ? 9 : 4;
    if (jb.wide) {
      // Few slabs and >= 128 input channels (the deep layers: most of the bytes).  Every warp transposes one output
      // channel x 128 input channels on its own: nine 512-byte rows in (one 16-byte load per lane and tap: whole DRAM
      // bursts instead of scattered 128-byte pieces), 1152 contiguous floats out, through its own slice of shared
      // memory with warp-level synchronisation only -- the eight warps of the CTA drift apart, so the SM always has
      // warps in their load phase.
      if (prev_shared) __syncthreads();
      prev_shared = false;
      const int wt = Cin >> 7;
      const int co = (lt / wt) * 8 + y, cw0 = (lt % wt) * 128;
      const float* s0 = jb.packed + (long long)co * 9 * Cin + cw0 + ci * 4;
      float4 acc[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sl = 0; sl < jb.slabs; ++sl) {
        const float* sp = s0 + (long long)sl * jb.slab_stride;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 v4 = __ldcg(reinterpret_cast<const float4*>(sp + (long long)t * Cin));
          acc[t].x += v4.x; acc[t].y += v4.y; acc[t].z += v4.z; acc[t].w += v4.w;
        }
      }
      float* Pw = P + y * (9 * UW_PITCH);
#pragma unroll
      for (int t = 0; t < 9; ++t) *reinterpret_cast<float4*>(Pw + t * UW_PITCH + ci * 4) = acc[t];
      __syncwarp();
      float* d0 = jb.dw + ((long long)co * Cin + cw0) * 9;
#pragma unroll 4
      for (int k = ci; k < 1152; k += 32) {
        const int i = k / 9, t = k - i * 9;
        const float v = Pw[t * UW_PITCH + i];
        d0[k] = accumulate ? d0[k] + v : v;
      }
      __syncwarp();
      continue;
    }
    prev_shared = true;
    __syncthreads();
    // The tile's source is 8 rows (co_t channels x `lanes` slab lanes) x T taps x 32 ci = 8 T x 8 float4 per slab round;
    // a thread owns up to three of them and reads 16 bytes per load: with 4-byte loads (one 128 B line per warp
    // request) the kernel stalled at 2.7 TB/s on the number of requests in flight, not on bytes.
    const int n4 = 64 * T;
    const float* src[3];
    int lane_of[3], prow[3];
    float4 acc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int idx = threadIdx.x + r * 256;
      acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      src[r] = nullptr;
      lane_of[r] = 0;
      prow[r] = 0;
      if (idx < n4) {
        const int yy = idx / (8 * T), rem = idx - yy * (8 * T);
        const int t = rem >> 3, c4 = rem & 7;
        const int co = co0 + yy / lanes;
        lane_of[r] = yy % lanes;
        prow[r] = (yy * 9 + t) * UP_PITCH + c4 * 4;
        src[r] = jb.packed + (jb.kind == 0 ? ((long long)co * 9 + t) * Cin : ((long long)t * Cout + co) * Cin) + ci0 + c4 * 4;
      }
    }
    const int rounds = (jb.slabs + lanes - 1) / lanes;
#pragma unroll 2
    for (int i = 0; i < rounds; ++i) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int sl = lane_of[r] + i * lanes;
        if (src[r] != nullptr && sl < jb.slabs) {
          const float4 v4 = __ldcg(reinterpret_cast<const float4*>(src[r] + (long long)sl * jb.slab_stride));
          acc[r].x += v4.x; acc[r].y += v4.y; acc[r].z += v4.z; acc[r].w += v4.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (src[r] != nullptr) {
        float* pr = P + prow[r];
        pr[0] = acc[r].x; pr[1] = acc[r].y; pr[2] = acc[r].z; pr[3] = acc[r].w;
      }
    }
    __syncthreads();
    if (jb.kind == 0) {
      for (int e = threadIdx.x; e < co_t * 288; e += 256) {     // torch [co][ci][t]: 288 contiguous floats per co
        const int c = e / 288, k = e - c * 288;
        const int i = k / 9, t = k - i * 9;
        float v = P[((c * lanes) * 9 + t) * UP_PITCH + i];
        for (int l = 1; l < lanes; ++l) v += P[((c * lanes + l) * 9 + t) * UP_PITCH + i];
        float* d = jb.dw + ((long long)(co0 + c) * Cin + ci0) * 9 + k;
        *d = accumulate ? *d + v : v;
      }
    } else {
      for (int e = threadIdx.x; e < 32 * co_t * 4; e += 256) {   // torch [ci][co][ab]: co_t * 4 contiguous floats per ci
        const int i = e / (co_t * 4), k = e - i * (co_t * 4);
        const int c = k >> 2, t = k & 3;
        float v = P[((c * lanes) * 9 + t) * UP_PITCH + i];
        for (int l = 1; l < lanes; ++l) v += P[((c * lanes + l) * 9 + t) * UP_PITCH + i];
        float* d = jb.dw + ((long long)(ci0 + i) * Cout + co0) * 4 + k;
        *d = accumulate ? *d + v : v;
      }
    }
  }
}

// Packed fp32 weight-gradient slabs (wgrad kernel layout) -> torch layout, one layer, plain gather (tests / tools).
//  mode 0 conv   : src[co][(r,s)][ci] -> dW[co][ci][r][s]
//  mode 2 upconv : src[(a,b,co)][ci]  -> dW[ci][co][a][b]
__global__ void __launch_bounds__(256) unpack_grad_kernel(const float* __restrict__ src, int slabs, long long slab_stride,
                                                          float* __restrict__ dw, int mode, int Cout, int Cin,
                                                          int accumulate) {
  const int T = (mode == 0) ? 9 : 4;
  const long long total = (long long)Cout * Cin * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long s;
    if (mode == 0) {
      const int t = (int)(i % 9);
      long long r = i / 9;
      const int ci = (int)(r % Cin);
      const int co = (int)(r / Cin);
      s = ((long long)co * 9 + t) * Cin + ci;
    } else {
      const int ab = (int)(i % 4);
      long long r = i / 4;
      const int co = (int)(r % Cout);
      const int ci = (int)(r / Cout);
      s = ((long long)ab * Cout + co) * Cin + ci;
    }
    float v = 0.0f;
    for (int k = 0; k < slabs; ++k) v += __ldg(src + (long long)k * slab_stride + s);
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// Partial rows -> vector, one producer, in row order (tests / stand-alone use of the bias-gradient producers):
// dst[e] (+)= sum_r part[r * stride + e], e < length.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int rows, long long stride,
                                                              int length, float* __restrict__ dst, int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= length) return;
  float v = 0.0f;
  for (int r = 0; r < rows; ++r) v += __ldg(part + (long long)r * stride + e);
  dst[e] = accumulate ? dst[e] + v : v;
}

// ---------------------------------------------------------------------------
// SGD with momentum over a flat fp32 arena (optim.SGD(lr, momentum), trainer.py:30,78):
//   buf = g (first step) | momentum*buf + g ;  p -= lr*buf.  No dampening/nesterov/weight decay.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ buf, long long n, float lr, float momentum,
                                                  int first_step) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 bv;
    if (first_step) {
      bv = gv;
    } else {
      bv = reinterpret_cast<float4*>(buf)[i];
      bv.x = fmaf(momentum, bv.x, gv.x); bv.y = fmaf(momentum, bv.y, gv.y);
      bv.z = fmaf(momentum, bv.z, gv.z); bv.w = fmaf(momentum, bv.w, gv.w);
    }
    pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
    reinterpret_cast<float4*>(buf)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  // tail
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gv = g[i];
    const float bv = first_step ? gv : fmaf(momentum, buf[i], gv);
    buf[i] = bv;
    p[i] -= lr * bv;
  }
}

#ifdef UB_DIAGNOSTICS
// Naive direct 3x3 valid conv on NHWC bf16 (fp32 accumulate): test-only comparator that
// shares nothing with the tcgen05 path.  One thread per output element.
__global__ void naive_conv3x3_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wpk,
                                     const float* __restrict__ bias, float* __restrict__ y, int n_img, int H, int W,
                                     int Cin, int Cout, int relu) {
  const int Ho = H - 2, Wo = W - 2;
  const long long total = (long long)n_img * Ho * Wo * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long pix = i / Cout;
    const int ox = (int)(pix % Wo);
    long long r2 = pix / Wo;
    const int oy = (int)(r2 % Ho);
    const int n = (int)(r2 / Ho);
    float a = bias ? bias[co] : 0.0f;
    for (int t = 0; t < 9; ++t) {
      const __nv_bfloat16* xp = x + (((long long)n * H + oy + t / 3) * W + ox + t % 3) * Cin;
      const __nv_bfloat16* wp = wpk + ((long long)co * 9 + t) * Cin;
      for (int c = 0; c < Cin; ++c) a = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), a);
    }
    y[i] = relu ? fmaxf(a, 0.0f) : a;
  }
}
#endif  // UB_DIAGNOSTICS

}  // namespace ub
