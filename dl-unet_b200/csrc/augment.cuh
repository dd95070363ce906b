// Training-sample augmentation on the device: the body of the reference's
// `ImageDataset.__getitem__` (data.py:93-137) and `elastic_transform` (data.py:225-245) for a whole batch:
//   crop window -> reflect pad by input_size (np.pad) -> scipy.ndimage.rotate (cubic spline, prefilter, reshape,
//   mode 'constant') + centre crop -> elastic deformation (Gaussian-filtered uniform displacement fields, bilinear
//   map_coordinates, mode 'constant') -> label crop + threshold, min-max normalisation of the image.
// The reference does this per sample on the CPU in about a second (scipy on a 1532 x 1532 padded image); here it is
// seven small kernels over the batch.  The arithmetic follows oracle/augment_oracle.py, which reproduces the
// reference bit for bit: all interpolation in fp64, uint8 outputs rounded half-up and clipped as scipy does.
//   * only a window of the padded image is materialised: the rotated crop reads within input_size/sqrt(2) + 2 pixels
//     of the centre, and the prefilter's pole (sqrt(3) - 2) forgets its start within ~30 pixels at fp64 precision;
//   * the spline prefilter runs one thread per line with the other axis coalesced: columns, transpose, columns.
// Everything here is latency / L2-bound integer and fp64 work on a few hundred MB.
#pragma once
#include <cstdint>

namespace ub {

struct AugGeom {          // per rotation angle (multiples of 30 degrees): output (i, j) reads R (t+i, l+j) + off
  double r00, r01, r10, r11, off0, off1;
  int t, l;
};
struct AugAngles {
  AugGeom g[12];
};

__device__ __forceinline__ int aug_reflect(int i, int pad, int n) {   // np.pad(mode='reflect'), any pad width
  const int p = 2 * (n - 1);
  int j = (i - pad) % p;
  if (j < 0) j += p;
  return j >= n ? p - j : j;
}

// ---- padded window of the crop, as fp64: A[b][arr][wy][wx], arr 0 = image, 1 = target
__global__ void __launch_bounds__(256) aug_window_kernel(const uint8_t* __restrict__ images,
                                                         const uint8_t* __restrict__ targets,
                                                         const int* __restrict__ sample, double* __restrict__ A,
                                                         int img_h, int img_w, int crop, int pad, int win0, int wn) {
  const int b = blockIdx.z >> 1, arr = blockIdx.z & 1;
  const int wy = blockIdx.y;
  const int* sp = sample + b * 4;
  const uint8_t* src = (arr ? targets : images) + (long long)sp[0] * img_h * img_w;
  const int sy = sp[1] + aug_reflect(win0 + wy, pad, crop);          // data.py:105-106: image[x:x+crop, y:y+crop]
  double* dst = A + (((long long)b * 2 + arr) * wn + wy) * wn;
  for (int wx = blockIdx.x * blockDim.x + threadIdx.x; wx < wn; wx += gridDim.x * blockDim.x)
    dst[wx] = (double)src[(long long)sy * img_w + sp[2] + aug_reflect(win0 + wx, pad, crop)];
}

// ---- cubic B-spline prefilter along dim 0 of [planes][n][n], one thread per column (scipy spline_filter1d, mirror
// boundary: what mode='constant' uses).  In place.
__global__ void __launch_bounds__(128) aug_prefilter_kernel(double* __restrict__ A, int n, int planes) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int pl = blockIdx.y;
  if (j >= n) return;
  double* a = A + (long long)pl * n * n + j;
  const double z = -0.26794919243112270647;   // sqrt(3) - 2
  const double gain = (1.0 - z) * (1.0 - 1.0 / z);
  double zk = 1.0, c0 = 0.0;
  const int horizon = n < 30 ? n : 30;         // z^30 ~ 7e-18
  for (int k = 0; k < horizon; ++k) {
    c0 += zk * (a[(long long)k * n] * gain);
    zk *= z;
  }
  // the recursions are sequential but their loads are not: fetch 8 rows ahead of the dependent chain
  double prev = c0;
  a[0] = c0;
  int k = 1;
  for (; k + 8 <= n; k += 8) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a[(long long)(k + i) * n];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      prev = v[i] * gain + z * prev;
      v[i] = prev;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[(long long)(k + i) * n] = v[i];
  }
  for (; k < n; ++k) {
    prev = a[(long long)k * n] * gain + z * prev;
    a[(long long)k * n] = prev;
  }
  double next = (z / (z * z - 1.0)) * (prev + z * a[(long long)(n - 2) * n]);
  a[(long long)(n - 1) * n] = next;
  k = n - 2;
  for (; k - 7 >= 0; k -= 8) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a[(long long)(k - i) * n];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      next = z * (next - v[i]);
      v[i] = next;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[(long long)(k - i) * n] = v[i];
  }
  for (; k >= 0; --k) {
    next = z * (next - a[(long long)k * n]);
    a[(long long)k * n] = next;
  }
}

__global__ void __launch_bounds__(256) aug_transpose_kernel(const double* __restrict__ A, double* __restrict__ T, int n) {
  __shared__ double tile[32][33];
  const long long base = (long long)blockIdx.z * n * n;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (y0 + r < n && x0 + tx < n) tile[r][tx] = A[base + (long long)(y0 + r) * n + x0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (x0 + r < n && y0 + tx < n) T[base + (long long)(x0 + r) * n + y0 + tx] = tile[tx][r];
}

__device__ __forceinline__ uint8_t aug_to_u8(double v) {   // scipy's float -> uint8: round half up, clip
  v = floor(v + 0.5);
  return (uint8_t)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));
}

__device__ __forceinline__ void aug_bspline3(double t, double (&w)[4]) {
  const double u = 1.0 - t;
  w[0] = u * u * u / 6.0;
  w[1] = (3.0 * t * t * t - 6.0 * t * t + 4.0) / 6.0;
  w[2] = (-3.0 * t * t * t + 3.0 * t * t + 3.0 * t + 1.0) / 6.0;
  w[3] = t * t * t / 6.0;
}

// ---- rotation + centre crop: Rm[b][arr][S][S] (uint8) from the TRANSPOSED coefficient window T[b][arr][x][y]
__global__ void __launch_bounds__(256) aug_rotate_kernel(const double* __restrict__ T, const int* __restrict__ sample,
                                                         uint8_t* __restrict__ Rm, AugAngles ang, int S, int win0,
                                                         int wn) {
  const int b = blockIdx.z >> 1, arr = blockIdx.z & 1;
  const int i = blockIdx.y;
  const AugGeom g = ang.g[(sample[b * 4 + 3] / 30) % 12];
  const double* c = T + ((long long)b * 2 + arr) * wn * wn;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < S; j += gridDim.x * blockDim.x) {
    const double oy = (double)(g.t + i), ox = (double)(g.l + j);
    const double yy = g.r00 * oy + g.r01 * ox + g.off0 - (double)win0;
    const double xx = g.r10 * oy + g.r11 * ox + g.off1 - (double)win0;
    const double fy = floor(yy), fx = floor(xx);
    double wy[4], wx[4];
    aug_bspline3(yy - fy, wy);
    aug_bspline3(xx - fx, wx);
    const int y0 = (int)fy - 1, x0 = (int)fx - 1;
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int yc = min(max(y0 + p, 0), wn - 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int xc = min(max(x0 + q, 0), wn - 1);
        acc += wy[p] * wx[q] * c[(long long)xc * wn + yc];
      }
    }
    Rm[(((long long)b * 2 + arr) * S + i) * S + j] = aug_to_u8(acc);
  }
}

// ---- Gaussian filter with zero boundary along one axis of [planes][S][S] (scipy gaussian_filter1d, mode
// 'constant'); `affine` applies in = 2 * in - 1 on the fly (first pass), `scale` multiplies the result (alpha).
// Block = 32 lines x 64 positions along the filtered axis, staged with its halo in shared memory (fp64); a thread
// produces 8 consecutive outputs from one sliding read of the taps (radius <= AUG_MAX_RADIUS).
constexpr int AUG_MAX_RADIUS = 48, AUG_BLUR_LEN = 64, AUG_BLUR_LINES = 32;
__global__ void __launch_bounds__(256) aug_blur_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                       const double* __restrict__ kern, int radius, int S, int axis,
                                                       int affine, double scale) {
  __shared__ double tile[AUG_BLUR_LEN + 2 * AUG_MAX_RADIUS][AUG_BLUR_LINES + 1];
  __shared__ double sk[2 * AUG_MAX_RADIUS + 1];
  const int pl = blockIdx.z;
  const int p0 = blockIdx.y * AUG_BLUR_LEN;     // first position along the filtered axis
  const int l0 = blockIdx.x * AUG_BLUR_LINES;   // first line (the other axis)
  const double* src = in + (long long)pl * S * S;
  const int span = AUG_BLUR_LEN + 2 * radius;
  for (int i = threadIdx.x; i < 2 * radius + 1; i += 256) sk[i] = kern[i];
  // stage: tile[pos][line]; consecutive threads read consecutive addresses of whichever axis is contiguous
  for (int i = threadIdx.x; i < span * AUG_BLUR_LINES; i += 256) {
    int pos, line;
    if (axis == 0) {   // filtered axis = rows: lines are columns (contiguous)
      line = i % AUG_BLUR_LINES;
      pos = i / AUG_BLUR_LINES;
    } else {           // filtered axis = columns (contiguous)
      pos = i % span;
      line = i / span;
    }
    const int p = p0 - radius + pos, l = l0 + line;
    double v = 0.0;
    if (p >= 0 && p < S && l < S) {
      v = axis == 0 ? src[(long long)p * S + l] : src[(long long)l * S + p];
      if (affine) v = 2.0 * v - 1.0;
    }
    tile[pos][line] = v;
  }
  __syncthreads();
  const int line = threadIdx.x & 31, seg = threadIdx.x >> 5;   // 8 segments of 8 outputs
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  const int base = seg * 8;
  for (int j = 0; j < 2 * radius + 8; ++j) {
    const double v = tile[base + j][line];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = j - i;                       // tap index for output i
      if (t >= 0 && t <= 2 * radius) acc[i] += sk[t] * v;
    }
  }
  const int l = l0 + line;
  if (l < S) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + base + i;
      if (p < S) out[(long long)pl * S * S + (axis == 0 ? (long long)p * S + l : (long long)l * S + p)] = acc[i] * scale;
    }
  }
}

__device__ __forceinline__ double aug_bilinear(const uint8_t* __restrict__ img, int S, double yy, double xx) {
  // map_coordinates(order=1, mode='constant', cval=0): outside [0, S-1] -> 0
  if (!(yy >= 0.0 && yy <= (double)(S - 1) && xx >= 0.0 && xx <= (double)(S - 1))) return 0.0;
  const double fy = floor(yy), fx = floor(xx);
  const double ty = yy - fy, tx = xx - fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, S - 1), x1 = min(x0 + 1, S - 1);
  double v = (1.0 - ty) * (1.0 - tx) * (double)img[(long long)y0 * S + x0];
  v += (1.0 - ty) * tx * (double)img[(long long)y0 * S + x1];
  v += ty * (1.0 - tx) * (double)img[(long long)y1 * S + x0];
  v += ty * tx * (double)img[(long long)y1 * S + x1];
  return v;
}

// ---- elastic deformation of image and target, label crop + threshold, per-sample min / max of the image
__global__ void __launch_bounds__(256) aug_elastic_kernel(const uint8_t* __restrict__ Rm, const double* __restrict__ disp,
                                                          uint8_t* __restrict__ E, long long* __restrict__ label,
                                                          int* __restrict__ minmax, int S, int crop, int lpad) {
  const int b = blockIdx.z, y = blockIdx.y;
  const uint8_t* img = Rm + ((long long)b * 2) * S * S;
  const uint8_t* tgt = img + (long long)S * S;
  const double* dx = disp + ((long long)b * 2) * S * S;   // axis-0 displacement (data.py:236, 241)
  const double* dy = dx + (long long)S * S;
  int lo = 255, hi = 0;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < S; x += gridDim.x * blockDim.x) {
    const long long p = (long long)y * S + x;
    const double yy = (double)y + dx[p], xx = (double)x + dy[p];
    const int v = aug_to_u8(aug_bilinear(img, S, yy, xx));
    E[(long long)b * S * S + p] = (uint8_t)v;
    lo = min(lo, v);
    hi = max(hi, v);
    const int ly = y - lpad, lx = x - lpad;
    if (ly >= 0 && ly < crop && lx >= 0 && lx < crop)     // gt[pad:original+pad, ...] > 127 (data.py:131-133)
      label[((long long)b * crop + ly) * crop + lx] = aug_to_u8(aug_bilinear(tgt, S, yy, xx)) > 127 ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(minmax + 2 * b, lo);
    atomicMax(minmax + 2 * b + 1, hi);
  }
}

// inp = (inp - min) / ptp in fp64, stored as fp32 (data.py:134, 136); a constant image gives 0/0 = nan as numpy does
__global__ void __launch_bounds__(256) aug_normalise_kernel(const uint8_t* __restrict__ E, const int* __restrict__ minmax,
                                                            float* __restrict__ out, long long per_sample,
                                                            long long total) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / per_sample);
    const double lo = (double)minmax[2 * b], ptp = (double)(minmax[2 * b + 1] - minmax[2 * b]);
    out[g] = (float)(((double)E[g] - lo) / ptp);
  }
}

__global__ void aug_init_minmax_kernel(int* __restrict__ minmax, int b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b) {
    minmax[2 * i] = 255;
    minmax[2 * i + 1] = 0;
  }
}

}  // namespace ub
