// Loss weight maps on the device (functions.py:7-78 `weighted_map`, functions.py:82-117 `class_balance`):
//   w = w_c + w0 * exp(-(d1 + d2)^2 / (2 sigma^2)) * (1 - gt)
// with d1 / d2 the exact Euclidean distances to the nearest / second-nearest 4-connected foreground object.
// The reference does this per sample on the CPU (OpenCV connectedComponents + one distanceTransform PER OBJECT +
// a sort over objects); here it is four integer kernels and the result is bit-compatible with the reference:
//   1. wm_init / wm_merge / wm_flatten : 4-connected component labelling.  One warp per row points every
//      foreground pixel at the start of its horizontal run (ballot + clz), vertical unions are issued only where a
//      new overlap between two runs begins (union-find with atomicMin), then every pixel is pointed at its root
//      (label = smallest linear pixel index of the component) and the roots are counted.
//   2. wm_columns : per column and row, the two nearest foreground pixels of DIFFERENT objects in that column
//      (dy, label), packed into 32 bits each.  Two per column are sufficient for the two nearest distinct objects
//      of any pixel: all candidates of a column share dx.
//   3. wm_rows : per pixel, scan the row's columns outwards (x, x-1, x+1, ...) keeping (d1^2, label1, d2^2) in
//      exact integer arithmetic, and stop as soon as dx^2 >= d2^2, or once both distances are known to be so
//      large that the exponential term is exactly +0 in fp32.
//      d = sqrtf((float)d^2) is what OpenCV's DIST_L2 / DIST_MASK_PRECISE transform returns (up to OpenCV's own
//      1-ulp sqrt error); the rest of the formula runs in fp64 and is rounded to fp32 once, like the numpy
//      expression it replaces.
// All byte / integer work, HBM- and latency-bound; one launch each for the whole batch.
#pragma once
#include <cstdint>

namespace ub {

constexpr uint32_t WM_INVALID = 0xffffffffu;
constexpr int WM_DY_BITS = 11;   // dy <= 2047, label + 1 < 2^21

template <typename LabelT>
__device__ __forceinline__ bool wm_fg(const LabelT* __restrict__ gt, long long i) {
  return (uint8_t)gt[i] != 0;   // connectedComponents runs on gt.astype(uint8): non-zero byte = foreground
}

// ---- counts: c1[b] = number of label == 1 ... (non-zero) pixels
template <typename LabelT>
__global__ void __launch_bounds__(256) wm_count_kernel(const LabelT* __restrict__ gt, int* __restrict__ count1,
                                                       int hw) {
  const int b = blockIdx.y;
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
    c += gt[(long long)b * hw + i] != 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count1 + b, c);
}

// ---- connected components (4-connectivity)
__device__ __forceinline__ int wm_find(const int* L, int i) {
  int p = __ldcg(L + i);   // L2 reads: other threads re-parent nodes with atomics while we walk
  while (p != i) {
    i = p;
    p = __ldcg(L + i);
  }
  return i;
}

__device__ __forceinline__ void wm_union(int* L, int a, int b) {
  while (true) {
    a = wm_find(L, a);
    b = wm_find(L, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(L + b, a);   // hang the larger root under the smaller one
    if (old == b) return;
    b = old;                               // somebody re-parented b meanwhile: retry from there
  }
}

// One warp per image row: every foreground pixel starts out pointing at the first pixel of its horizontal run
// (ballot + count-leading-zeros per 32-pixel chunk, run start carried across chunks), so the horizontal unions
// are done before the union-find starts.
template <typename LabelT>
__global__ void __launch_bounds__(256) wm_init_kernel(const LabelT* __restrict__ gt, int* __restrict__ L, int H,
                                                      int W, int B) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * H) return;
  const int y = (int)(row % H);
  const LabelT* g = gt + row * W;
  int* Lr = L + row * W;
  int carry = -1;   // run start (x) if the pixel just left of this chunk is foreground
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const bool fg = x < W && (uint8_t)g[x] != 0;   // connectedComponents runs on gt.astype(uint8)
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    const uint32_t zeros_below = ~m & ((1u << lane) - 1u);
    int start;
    if (zeros_below) start = x0 + (32 - __clz(zeros_below));   // just right of the nearest background pixel
    else start = carry >= 0 ? carry : x0;
    if (x < W) Lr[x] = fg ? y * W + start : -1;
    // carry for the next chunk: run start of lane 31 if it is foreground
    const int s31 = __shfl_sync(0xffffffffu, start, 31);
    carry = (m >> 31) ? s31 : -1;
  }
}

// Vertical unions, only where a new overlap between two runs begins: pixel and its upper neighbour are foreground
// and (left, upper-left) are not both foreground (otherwise the pixel to the left already joined the two runs).
__global__ void __launch_bounds__(256) wm_merge_kernel(int* L, int H, int W, int B) {
  const int hw = H * W;
  const long long total = (long long)B * hw;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / hw), i = (int)(g - (long long)b * hw);
    int* Lb = L + (long long)b * hw;
    if (i < W || Lb[i] < 0 || Lb[i - W] < 0) continue;
    const int x = i % W;
    if (x > 0 && Lb[i - 1] >= 0 && Lb[i - W - 1] >= 0) continue;
    wm_union(Lb, i, i - W);
  }
}

__global__ void __launch_bounds__(256) wm_flatten_kernel(int* L, int* __restrict__ n_obj, int hw, long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long base = g - g % hw;
    const int i = (int)(g - base);
    if (L[g] < 0) continue;
    const int r = wm_find(L + base, i);   // roots are fixed points, so concurrent flattening is safe
    L[g] = r;
    if (r == i) atomicAdd(n_obj + g / hw, 1);   // one root per object
  }
}

// ---- column pass: thread = (sample, 32-row segment, column) -> cand[b][y][2][x]: the two nearest foreground pixels
// OF DIFFERENT OBJECTS in column x as seen from row y, packed (label + 1) << 11 | dy, nearest first.
// Two per column are sufficient for the two nearest distinct objects of any pixel of the row: all candidates of one
// column share dx, so a third object of that column is beaten by the two before it, and a farther pixel of the same
// object by the nearer one.  They are the best two of four: the nearest pixel above / below and the nearest pixel
// above / below that belongs to another object than that one (found by walking the column down, then up).
// Distances of `cap` rows or more are dropped: beyond d1 + d2 >= cap the weight term is exactly +0 in fp32 (see
// wm_rows_kernel), so the state a segment inherits from the rows before it is found by looking back < cap rows.
constexpr int WM_SEG = 32;
__device__ __forceinline__ uint32_t wm_pack(int dy, int label) { return ((uint32_t)(label + 1) << WM_DY_BITS) | (uint32_t)dy; }
__device__ __forceinline__ uint32_t wm_dy(uint32_t c) { return c & ((1u << WM_DY_BITS) - 1u); }
__device__ __forceinline__ uint32_t wm_label(uint32_t c) { return c >> WM_DY_BITS; }

struct WmScan {
  int y1, l1, y2, l2;   // row / label of the nearest foreground pixel, and of the nearest one of another object
  __device__ __forceinline__ void see(int y, int l) {      // rows arrive in scan order (monotone y)
    if (l < 0) return;
    if (l != l1) {
      y2 = y1;
      l2 = l1;
      l1 = l;
    }
    y1 = y;
  }
  // state before a segment: walk away from it (step = -1 above, +1 below) until two objects are known
  __device__ __forceinline__ void inherit(const int* __restrict__ col, int W, int y_from, int y_stop, int step) {
    y1 = l1 = y2 = l2 = -1;
    for (int y = y_from; y != y_stop; y += step) {
      const int l = col[(long long)y * W];
      if (l < 0) continue;
      if (l1 < 0) {
        y1 = y;
        l1 = l;
      } else if (l != l1) {
        y2 = y;
        l2 = l;
        return;
      }
    }
  }
};

// best two of (a1, a2, b1, b2) with different labels; a1 / b1 are the nearest of their direction
__device__ __forceinline__ void wm_best2(uint32_t a1, uint32_t a2, uint32_t b1, uint32_t b2, uint32_t& c1, uint32_t& c2) {
  // WM_INVALID has dy = 2047 = the largest, so plain comparisons of dy sort it last
  c1 = (wm_dy(b1) < wm_dy(a1)) ? b1 : a1;
  c2 = WM_INVALID;
  if (c1 == WM_INVALID) return;
  const uint32_t l = wm_label(c1);
  uint32_t best = 1u << WM_DY_BITS;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t c = i == 0 ? a1 : i == 1 ? a2 : i == 2 ? b1 : b2;
    if (c != WM_INVALID && wm_label(c) != l && wm_dy(c) < best) {
      best = wm_dy(c);
      c2 = c;
    }
  }
}

__global__ void __launch_bounds__(128) wm_columns_kernel(const int* __restrict__ L, uint32_t* __restrict__ cand, int H,
                                                         int W, int cap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * WM_SEG, b = blockIdx.z;
  if (x >= W) return;
  const int y_end = min(H, y0 + WM_SEG);
  const int* Lb = L + (long long)b * H * W + x;
  uint32_t* cb = cand + (long long)b * H * 2 * W + x;
  WmScan s;
  s.inherit(Lb, W, y0 - 1, max(-1, y0 - cap), -1);
  for (int y = y0; y < y_end; ++y) {          // candidates above: parked in the output planes until the walk back
    s.see(y, Lb[(long long)y * W]);
    const int d1 = y - s.y1, d2 = y - s.y2;
    cb[((long long)y * 2 + 0) * W] = (s.l1 >= 0 && d1 < cap) ? wm_pack(d1, s.l1) : WM_INVALID;
    cb[((long long)y * 2 + 1) * W] = (s.l2 >= 0 && d2 < cap) ? wm_pack(d2, s.l2) : WM_INVALID;
  }
  s.inherit(Lb, W, y_end, min(H, y_end - 1 + cap), 1);
  for (int y = y_end - 1; y >= y0; --y) {     // candidates below, merged with the ones above
    s.see(y, Lb[(long long)y * W]);
    const int d1 = s.y1 - y, d2 = s.y2 - y;
    const uint32_t b1 = (s.l1 >= 0 && d1 < cap) ? wm_pack(d1, s.l1) : WM_INVALID;
    const uint32_t b2 = (s.l2 >= 0 && d2 < cap) ? wm_pack(d2, s.l2) : WM_INVALID;
    uint32_t c1, c2;
    wm_best2(cb[((long long)y * 2 + 0) * W], cb[((long long)y * 2 + 1) * W], b1, b2, c1, c2);
    cb[((long long)y * 2 + 0) * W] = c1;
    cb[((long long)y * 2 + 1) * W] = c2;
  }
}

// ---- row pass + formula.  grid = (ceil(W / 128), H, B); the row's candidates live in shared memory.
struct WmState {
  uint32_t d1, d2;   // squared distances (exact)
  uint32_t l1;       // label + 1 of the nearest object
};

__device__ __forceinline__ void wm_offer(WmState& s, uint32_t l, uint32_t d) {
  if (l == s.l1) {
    s.d1 = min(s.d1, d);
  } else if (d < s.d1) {   // a new nearest object: the old nearest becomes the best of "all the others"
    s.d2 = s.d1;
    s.d1 = d;
    s.l1 = l;
  } else {
    s.d2 = min(s.d2, d);
  }
}

template <typename LabelT>
__global__ void __launch_bounds__(128) wm_rows_kernel(const LabelT* __restrict__ gt, const uint32_t* __restrict__ cand,
                                                      const int* __restrict__ count1, const int* __restrict__ n_obj,
                                                      float* __restrict__ out, int* __restrict__ status, int H, int W,
                                                      double w0, double two_sigma2, int truncate_wc, int cap) {
  extern __shared__ uint32_t srow[];   // [2][W]: the two candidates of every column of this row
  const int y = blockIdx.y, b = blockIdx.z;
  const uint32_t* crow = cand + ((long long)b * H + y) * 2 * W;
  for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) srow[i] = crow[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const long long pix = ((long long)b * H + y) * W + x;
  const bool fg = gt[pix] != 0;
  WmState s{WM_INVALID, WM_INVALID, 0u};
  // `cap`: beyond d1 + d2 >= cap the term w0 * exp(-(d1+d2)^2 / 2 sigma^2) is below 2^-150 and rounds to +0 in
  // fp32 whatever w_c is, so neither distance needs to be known beyond it (set by the host from w0 and sigma).
  const bool one_object = n_obj[b] == 1;   // then d2 = 0 (functions.py:62-64)
  const int kmax = min(W, cap);
  const uint32_t cap2 = (uint32_t)cap * (uint32_t)cap;
  for (int k = 0; k < kmax && !fg; ++k) {   // foreground pixels get w_d = 0: (1 - gt) in functions.py:70
    const uint32_t k2 = (uint32_t)k * (uint32_t)k;
    if (k2 >= s.d2) break;   // nothing at |dx| >= k can improve either distance
    // every unseen candidate is >= k away, so d1_final >= min(d1, k) and d2_final >= min(d2, k): once both are
    // >= k >= cap / 2 the sum is >= cap and the weight term is +0 whatever is still out there
    if (!one_object && 4u * k2 >= cap2 && s.d1 >= k2) break;
    const int xl = x - k, xr = x + k;
    if (xl < 0 && xr >= W) break;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int xc = side == 0 ? xl : xr;
      if (xc < 0 || xc >= W || (side == 1 && k == 0)) continue;
      const uint32_t c1 = srow[xc];
      if (c1 == WM_INVALID) continue;
      const uint32_t dy1 = wm_dy(c1), e1 = k2 + dy1 * dy1;
      if (e1 >= s.d2) continue;                 // the column's nearest cannot improve anything, nor can its second
      wm_offer(s, wm_label(c1), e1);
      const uint32_t c2 = srow[W + xc];
      if (c2 == WM_INVALID) continue;
      const uint32_t dy2 = wm_dy(c2);
      wm_offer(s, wm_label(c2), k2 + dy2 * dy2);
    }
  }
  const int hw = H * W;
  const int c1 = count1[b], c0 = hw - c1;
  // w_c (functions.py:34-40): counts[1] / counts[class]; held in the label's dtype, so an integer label truncates
  float wc = fg ? 1.0f : (float)c1 / (float)c0;
  if (truncate_wc) wc = truncf(wc);
  double wd = 0.0;
  if (!fg && s.d1 != WM_INVALID && (one_object || s.d2 != WM_INVALID)) {
    const double d1 = (double)sqrtf((float)s.d1);
    const double d2 = one_object ? 0.0 : (double)sqrtf((float)s.d2);
    const double t = d1 + d2;
    wd = w0 * exp(-(t * t) / two_sigma2);   // a distance that was cut off at `cap` only ever lands here as exp -> 0
  }
  out[pix] = (float)((double)wc + wd);
  if (x == 0 && y == 0) {
    int st = 0;
    if (c1 == 0 || c0 == 0) st |= 1;   // a single class: the reference's counts[1] raises IndexError
    if (c1 == 0) st |= 2;              // no object: maps[0] raises IndexError
    status[b] = st;
  }
}

// class_balance (functions.py:82-117): w = counts[1] / counts[class] in fp32
template <typename LabelT>
__global__ void __launch_bounds__(256) wm_class_balance_kernel(const LabelT* __restrict__ gt,
                                                               const int* __restrict__ count1, float* __restrict__ out,
                                                               int* __restrict__ status, int hw) {
  const int b = blockIdx.y;
  const int c1 = count1[b], c0 = hw - c1;
  const float w0 = (float)c1 / (float)c0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
    out[(long long)b * hw + i] = gt[(long long)b * hw + i] != 0 ? 1.0f : w0;
  if (blockIdx.x == 0 && threadIdx.x == 0) status[b] = (c1 == 0 || c0 == 0) ? 1 : 0;
}

// ---- mask metrics (functions.py:150-213): counts for IoU = |p & l| / |p | l| and pixel error = sum|p - l| / n.
// out[0] += |p & l|, out[1] += |p | l|, out[2] += sum |p - l| (64-bit counters, caller zeroes).
template <typename PT, typename LT>
__global__ void __launch_bounds__(256) mask_metrics_kernel(const PT* __restrict__ pred, const LT* __restrict__ label,
                                                           unsigned long long* __restrict__ out, long long n) {
  unsigned int inter = 0, uni = 0;
  unsigned long long diff = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long p = (long long)pred[i], l = (long long)label[i];
    inter += (p != 0 && l != 0) ? 1u : 0u;      // np.logical_and / np.logical_or
    uni += (p != 0 || l != 0) ? 1u : 0u;
    diff += (unsigned long long)(p > l ? p - l : l - p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
    diff += __shfl_xor_sync(0xffffffffu, diff, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (inter) atomicAdd(out, (unsigned long long)inter);
    if (uni) atomicAdd(out + 1, (unsigned long long)uni);
    if (diff) atomicAdd(out + 2, diff);
  }
}

// ---- mirror padding of a test image up to the network's input size (data.py:249-312 `mirror_transform`):
// out[b, y, x] = in[b, m(y), m(x)],  m(i) = pad - i below the image (reflect, edge not repeated) and
// 2n - 1 - (i - pad) above it (the reference's slices repeat the edge row / column on that side).
__device__ __forceinline__ int mirror_index(int i, int pad, int n) {
  const int j = i - pad;
  return j < 0 ? -j : (j >= n ? 2 * n - 1 - j : j);
}

__global__ void __launch_bounds__(256) mirror_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                                                         int size, int pad, long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(g % size);
    const long long r = g / size;
    const int y = (int)(r % size);
    const long long b = r / size;
    out[g] = __ldg(in + (b * n + mirror_index(y, pad, n)) * n + mirror_index(x, pad, n));
  }
}

// ---- overlap-tile inference (Ronneberger's tiling strategy): input tiles gathered straight from
// the image and output mask tiles scattered straight into the stitched mask, so neither the reflect-padded image nor
// the full tile stack ever exists.  Tile t = (ty, tx) of an nx-wide grid covers output pixels [ty*To, (ty+1)*To) x
// [tx*To, ...) and reads input pixels shifted by -halo: coordinates below 0 reflect without repeating the edge
// (F.pad "reflect"), coordinates beyond the reflect-padded extent H + halo repeat its last row / column (the
// right / bottom fill up to a whole number of tiles), exactly as b200.tiling.cut_tiles' F.pad + unfold did.
__device__ __forceinline__ int tile_src_index(int i, int n, int halo) {   // i: coordinate in the image frame
  if (i > n - 1 + halo) i = n - 1 + halo;
  if (i < 0) i = -i;
  if (i > n - 1) i = 2 * (n - 1) - i;
  return i;
}

// image: rows [row0, row0 + rows) of an H x W fp32 image (a rank may hold only the band its tiles touch)
__global__ void __launch_bounds__(256) cut_tiles_kernel(const float* __restrict__ image, int H, int W, int row0,
                                                        const int* __restrict__ tile_ids, int nx, int tile_in,
                                                        int tile_out, float* __restrict__ tiles, long long total) {
  const int halo = (tile_in - tile_out) / 2;
  const long long plane = (long long)tile_in * tile_in;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(g / plane);
    const int r = (int)(g - (long long)k * plane);
    const int py = r / tile_in, px = r - py * tile_in;
    const int t = __ldg(tile_ids + k);
    const int ty = t / nx, tx = t - ty * nx;
    const int y = tile_src_index(ty * tile_out + py - halo, H, halo);
    const int x = tile_src_index(tx * tile_out + px - halo, W, halo);
    tiles[g] = __ldg(image + (long long)(y - row0) * W + x);
  }
}

__global__ void __launch_bounds__(256) stitch_tiles_kernel(const unsigned char* __restrict__ masks,
                                                           const int* __restrict__ tile_ids, int nx, int tile_out,
                                                           unsigned char* __restrict__ out, int H, int W, long long total) {
  const long long plane = (long long)tile_out * tile_out;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(g / plane);
    const int r = (int)(g - (long long)k * plane);
    const int py = r / tile_out, px = r - py * tile_out;
    const int t = __ldg(tile_ids + k);
    const int ty = t / nx, tx = t - ty * nx;
    const int y = ty * tile_out + py, x = tx * tile_out + px;
    if (y < H && x < W) out[(long long)y * W + x] = masks[g];
  }
}

}  // namespace ub
