/*
 * unet_b200.h -- C ABI of the B200-native U-Net hot path.
 *
 * The reference (nsirons/DL-unet) has no FFI of its own: its operator layer is
 * whatever ATen kernel `torch.nn` dispatches to.  Each entry point below replaces
 * one of those dispatches; the reference call site it stands in for is cited per
 * function (paths relative to the reference repo).  INTEGRATION.md shows the
 * ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer unless stated otherwise; nothing is
 *     allocated or freed by the library; `stream` is a cudaStream_t cast to void*.
 *   - activations are NHWC bf16, described by ub_tensor4 (channel stride is 1;
 *     n/y/x strides are in ELEMENTS, so channel-slices of a wider buffer and
 *     windows of a larger image are expressible -- this is how the skip
 *     connection is written in place).
 *   - return value: 0 = ok, negative = UB_E_* (bad shape / alignment / no
 *     driver), positive = cudaError_t of the failed launch.  No entry point
 *     synchronises the device, exits the process or falls back to the CPU.
 *   - all entry points are re-entrant and capturable into a CUDA graph.
 */
#ifndef UNET_B200_H_
#define UNET_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define UB_E_SHAPE (-1)     /* dimension not supported (e.g. channels not a multiple of 64) */
#define UB_E_ALIGN (-2)     /* pointer / stride not 16-byte aligned */
#define UB_E_DRIVER (-3)    /* cuTensorMapEncodeTiled unavailable (no CUDA driver) */
#define UB_E_TENSORMAP (-4) /* the driver rejected a tensor map */
#define UB_E_ARG (-5)       /* null pointer / bad enum */

typedef struct ub_tensor4 {
  void* ptr;             /* first element of the view */
  int n, h, w, c;        /* extents of the view */
  long long sn, sy, sx;  /* element strides of n, y (row), x (pixel); channel stride is 1 */
} ub_tensor4;

int ub_version(void);
const char* ub_error_string(int code);
/* Non-zero after a kernel aborted on a pipeline dead-lock guard; host memory out-param. Synchronises. */
int ub_device_error(unsigned int* out_code);
/* experiment: device buffer of 148 x 8 uint64 that igemm2 launches fill with the clock sums of their epilogue phases
 * (wait for the MMAs, TMEM load + arithmetic + staging, store wait, chunk barrier, store issue, chunks, total); NULL = off.
 * The probes are compiled into the diagnostics build (libunet_b200_diag.so) only; the product library accepts and ignores the buffer. */
void ub_set_prof_buffer(void* dev_ptr);
/* Size the persistent grids of the following launches for n_sms SMs instead of all of them (0 = all): leaves SMs free
 * for a collective that runs concurrently on another stream (data-parallel gradient all-reduce), so that neither
 * kernel waits for the other's CTAs to drain.  Host-side state, read at launch time. */
void ub_set_sm_limit(int n_sms);
/* Tuning overrides for experiments: <=0 restores the heuristic. */
void ub_set_tuning(int force_bn, int force_bw, int force_bh, int force_splits);
/* Experiment switches (A/B measurements; the defaults are the product configuration).
 * key 1: 3x3 wgrad kernel (0 = persistent tap-per-unit kernel, 1 = haloed patch, 2 = haloed patch + fused s taps, 3 = auto
 *        [default], 4 = tap-row-per-CTA N=192, 5 = shifted-dY kernel for 64 output channels).
 * key 2: 3x3 fwd/dgrad kernel (0 = tap-per-load igemm only, 1 = haloed-patch kernel on wide shallow layers [default],
 *        2 = wherever legal).
 * key 5: igemm ablation (1 = do not wait for earlier TMA stores): timing experiments only.
 * key 7: host-side cache of encoded TMA tensor maps (1 = on [default], 0 = encode on every call).
 * key 8: conv11c weight gradient on the tensor pipe (1 = on [default], 0 = CUDA-core kernel).
 * key 11: N = 256 tiles of the tap-per-load conv / up-conv kernel on CTA pairs, tcgen05 cta_group::2 (1 = on [default],
 *         0 = one CTA per 128 x 256 tile).
 * key 13: stream-K in the CTA-pair kernel (0 = whole units round-robin, 1 = contiguous K-step ranges per pair when the idle
 *         tail of the last wave of units is worth >= 40 K steps [default], 2 = always).  Uses a library-owned 32 MB workspace per stream.
 * key 14: shared-memory split of the CTA-pair kernel (0 = by K-loop length [default], 1 = 5 operand stages + 4 output
 *         staging buffers, 2 = 3 + 8: up to 7 TMA stores in flight).
 * key 18: CTA-pair tap-per-load kernel, 3x3 convolutions: TMA boxes that also span several images so that small maps fill the
 *         128 MMA rows (1 = on [default], 0 = boxes inside one image).
 * key 19: weight gradient of the layers with 64 output channels: nine taps in ten accumulator slots (N = 192 + N = 128 per K
 *         step, one-pixel- and one-row-shifted dY copies; 1 = on [default]) or in twelve (three N = 128 steps; 0).
 * key 17: the haloed-patch kernel on CTA pairs (1 = on [default]: one M = 256 tcgen05.mma.cta_group::2 per step for two adjacent
 *         pixel tiles, each CTA keeps half of every weight slab; 0 = one CTA per tile). */
void ub_set_option(int key, int value);

/* ---- weight packing (derived caches of the fp32 masters; network.py:23-58) ----------
 * mode 0: Conv2d [Cout,Cin,3,3]           -> forward operand  [Cout][9][Cin]   bf16
 * mode 1: Conv2d [Cout,Cin,3,3]           -> dgrad operand    [Cin][9][Cout]   bf16 (taps rotated 180 deg)
 * mode 2: ConvTranspose2d [Cin,Cout,2,2]  -> forward operand  [4*Cout][Cin]    bf16 (rows ordered a,b,co)
 * mode 3: ConvTranspose2d [Cin,Cout,2,2]  -> dgrad operand    [Cin][4*Cout]    bf16                      */
int ub_pack_weight(const float* w, void* dst_bf16, int mode, int cout, int cin, void* stream);
/* inverse of modes 0 / 2 for the fp32 gradient slabs the wgrad kernels produce: dw (torch layout) = (dw +) the sum,
 * in slab order, of n_slabs packed partial gradients `slab_stride` elements apart (one layer, plain gather) */
int ub_unpack_grad(const float* slabs, int n_slabs, long long slab_stride, float* dw, int mode, int cout, int cin,
                   int accumulate, void* stream);

/* All conv / up-conv layers in one launch.  A job table (ub_job_table_bytes(n) bytes) is filled on the host by
 * ub_job_table_fill from per-layer pointer arrays and copied to the device by the caller.  It returns the number of
 * pack tiles (pass it as `total` to ub_pack_all; negative = error) and writes the n_jobs + 1 prefix sums of the
 * UNPACK tiles to utile_begin_out.  `packed[i]` is layer i's slab workspace and slabs[i] = ub_wgrad_slabs(...) for the
 * shape it is used with.  ub_pack_all writes both bf16 operands of every layer; ub_unpack_range(lo, hi) sums the
 * slabs of the layers whose unpack tiles lie in [lo, hi) in fixed order and writes the torch-layout gradients
 * (lo = utile_begin[i], hi = utile_begin[i + 1]: layer i alone; 0 .. utile_begin[n]: all layers). */
int ub_job_table_bytes(int n_jobs);
int ub_job_table_fill(void* table_host, int n_jobs, const void* const* w, void* const* dst_a, void* const* dst_b,
                      const void* const* packed, void* const* dw, const int* kind, const int* cout, const int* cin,
                      const int* slabs, int* utile_begin_out);
int ub_pack_all(const void* table_dev, int n_jobs, long long total, void* stream);
int ub_unpack_range(const void* table_dev, int n_jobs, int tile_lo, int tile_hi, int accumulate, void* stream);

/* ---- 3x3 valid convolution (nn.Conv2d(k=3) + F.relu; network.py:131-188) -------------
 * y[n,oy,ox,co] = act(bias[co] + sum_{r,s,ci} x[n,oy+r,ox+s,ci] * W[co,ci,r,s]);  y is (h-2)x(w-2). */
int ub_conv3x3_fwd(const ub_tensor4* x, const void* w_fwd, const float* bias, const ub_tensor4* y, int relu,
                   void* stream);
/* conv12e + finalconv in one launch (network.py:187-190; tester.py:30): the last 3x3 conv (<= 256 -> 64 channels,
 * bias + ReLU) with the 1x1 64 -> 2 head evaluated in the epilogue on the fp32 accumulators, i.e. BEFORE the bf16
 * rounding of the activation.  y may be NULL (inference: the 64-channel activation is never written to HBM).
 * logits: fp32 [n,2,h-2,w-2] NCHW or NULL; argmax: uint8 [n,h-2,w-2] or NULL (1 iff z1 > z0: ties -> class 0). */
int ub_conv3x3_fwd_head(const ub_tensor4* x, const void* w_fwd, const float* bias, const ub_tensor4* y,
                        const float* head_w, const float* head_b, float* logits_nchw, unsigned char* argmax,
                        void* stream);
/* conv{1,2}2c + max_pool2d in one launch (network.py:132-133, 138-139): a 3x3 conv (<= 256 -> cout in {64, 128, 192,
 * 256} channels, bias + ReLU) whose epilogue also pools every staged 8 x 16 pixel tile 2x2: `pooled` (dense
 * [n,(h-2)/2,(w-2)/2,cout] bf16), `idx` (1 byte per pooled element) and the skip copy into `cat` follow
 * ub_pool2x2_fwd's rules exactly.  y may be NULL: the full-resolution activation is then never written to HBM (the
 * backward pass does not need it: its ReLU mask is pooled > 0 and the layer's wgrad reads the layer's INPUT). */
int ub_conv3x3_fwd_pool(const ub_tensor4* x, const void* w_fwd, const float* bias, int cout, const ub_tensor4* y,
                        void* pooled_bf16, unsigned char* idx, const ub_tensor4* cat, int c_off, void* stream);
/* dx = full correlation of dy with the rotated weights; if relu_src != NULL the ReLU backward of the
 * layer that produced x is fused: dx *= (relu_src > 0).  (autograd of network.py:131-188) */
int ub_conv3x3_dgrad(const ub_tensor4* dy, const void* w_dgrad, const ub_tensor4* dx, const ub_tensor4* relu_src,
                     void* stream);
/* ---- partial sums (bias gradients and the two small weight gradients) ----------------------------------------
 * Every kernel that reduces over pixels into a short vector leaves PARTIAL ROWS instead of accumulating with atomics:
 * CTA r stores row r of part[rows][stride]; the rows are summed in row order by ub_unpack_range (vector jobs, kind 3,
 * of the job table) or ub_reduce_partials -- fixed summation order, bit-reproducible gradients, nothing to zero.
 * rows = ub_partial_rows(producer, ...): a function of the shape and the SM count only.
 *   producer 0: ub_pool2x2_bwd with dbias     (n, h, w, c of dx; c_dy, c0 unused)             stride = c
 *   producer 1: ub_colsum                     (n*h*w pixels, c channels)                       stride = c
 *   producer 2: ub_conv3x3_dgrad_bias         (n, h, w, c of dx; c_dy channels of dy; c0)      stride = c - c0
 *   producer 3: ub_head_bwd                   (n, h, w)                                        stride = 194
 *   producer 4: ub_conv1_wgrad                (n, h, w of the image)                           stride = 640
 *   producer 5: ub_upconv2x2_dgrad_bias       (n, h, w, c of dx)                               stride = c */
int ub_partial_rows(int producer, int n, int h, int w, int c, int c_dy, int c0);
/* dst[e] (+)= sum_r part[r * stride + e], e < length (one producer; the job table does all of them in one launch) */
int ub_reduce_partials(const float* part, int rows, long long stride, int length, float* dst, int accumulate, void* stream);

/* ub_conv3x3_dgrad, and partial rows of sum over pixels of dx[.., c] for the channels c >= c0 (row length c - c0): the
 * bias gradient of the layer whose output gradient dx is (c0 > 0: the up-conv half of a concat buffer).  Fused into
 * the epilogue where the chosen kernel supports it (register accumulators, one row per CTA), one extra column-sum
 * pass over dx otherwise. */
int ub_conv3x3_dgrad_bias(const ub_tensor4* dy, const void* w_dgrad, const ub_tensor4* dx, const ub_tensor4* relu_src,
                          float* dbias_part, int dbias_c0, void* stream);
/* Weight gradients are DETERMINISTIC: the reduction over pixels is split over CTAs (split-K), every split STORES
 * its fp32 partial sums into its own slab and the slabs are added in fixed order by ub_unpack_range / ub_unpack_grad
 * (no atomics anywhere; two runs give bit-identical gradients).
 * ub_wgrad_slabs: number of slabs the wgrad of this shape writes (kind 0: 3x3 conv, x is h_x x w_x, dy two smaller;
 * kind 2: 2x2 up-conv, dy twice as large); depends on the shape and the SM count only.  Negative = error.
 * dw_slabs: fp32 [slabs][Cout][9][Cin] (conv) or [slabs][(a,b,co)][Cin] (up-conv), fully overwritten. */
int ub_wgrad_slabs(int kind, int n, int h_x, int w_x, int cin, int cout);
int ub_conv3x3_wgrad(const ub_tensor4* x, const ub_tensor4* dy, float* dw_slabs, void* stream);

/* ---- 2x2 stride-2 up-convolution (nn.ConvTranspose2d(k=2,s=2); network.py:159,167,175,183) ----
 * y[n,2i+a,2j+b,co] = bias[co] + sum_ci x[n,i,j,ci] * W[ci,co,a,b]; y is normally the upper channel
 * half of the decoder's concat buffer (network.py:127), expressed through y->ptr / strides.          */
int ub_upconv2x2_fwd(const ub_tensor4* x, const void* w_fwd, const float* bias, const ub_tensor4* y, void* stream);
int ub_upconv2x2_dgrad(const ub_tensor4* dy, const void* w_dgrad, const ub_tensor4* dx, const ub_tensor4* relu_src,
                       void* stream);
/* ub_upconv2x2_dgrad, and partial rows (producer 5) of the column sums of dx: the bias gradient of the conv that produced
 * the up-convolution's input (fused into the epilogue on the CTA-pair kernel, one extra column-sum pass otherwise) */
int ub_upconv2x2_dgrad_bias(const ub_tensor4* dy, const void* w_dgrad, const ub_tensor4* dx, const ub_tensor4* relu_src,
                            float* dbias_part, void* stream);
int ub_upconv2x2_wgrad(const ub_tensor4* x, const ub_tensor4* dy, float* dw_slabs, void* stream);

/* ---- first convolution, Cin = 1 (network.py:131), fp32 NCHW image in, bf16 NHWC out, bias+ReLU fused ---- */
int ub_conv1_fwd(const float* image, const float* w, const float* bias, void* y_bf16, int n, int h, int w_,
                 void* stream);
/* partial rows [rows][640]: dW[64][9] then db[64] (producer 4).  ws_bf16 (ub_conv1_wgrad_workspace_bytes, 16-byte
 * aligned): scratch for a bf16 copy of the image -- with it the gradient runs on the tensor pipe (the image enters the
 * products rounded to bf16; exact for bf16-representable images); NULL or ub_set_option(8, 0): CUDA-core kernel on the
 * fp32 image.  The row count depends on which of the two runs (ub_partial_rows reports it for the current option). */
long long ub_conv1_wgrad_workspace_bytes(int n, int h, int w_);
int ub_conv1_wgrad(const float* image, const void* dy_bf16, float* part, void* ws_bf16, int n, int h, int w_, void* stream);

/* ---- 2x2 max pool (F.max_pool2d(t,2,2); network.py:133,139,145,151) + skip write (network.py:108-127) ----
 * x: dense [n,h,w,c] bf16.  y: dense [n,h/2,w/2,c].  idx: 1 byte per output element, code = dy*2+dx of the
 * first maximum in row-major window order (ATen's tie rule; NaN wins).  If cat != NULL the pooled value
 * is also written to cat[n, oy - c_off, ox - c_off, 0:c] where in range: c_off = int((Hpool - Hup)*0.5)
 * is crop_and_concat's offset (negative = zero padding, the border is never written).                */
int ub_pool2x2_fwd(const void* x, void* y, unsigned char* idx, const ub_tensor4* cat, int c_off, int n, int h, int w,
                   int c, void* stream);
/* dx[n,h,w,c] = scatter((dy + window(dcat)) * (pooled > 0)) by idx; dy / dcat may be NULL.
 * dbias (may be NULL): partial rows [rows][c] (producer 0) of the column sums of dx = the bias gradient of the conv
 * that fed the pool (saves the separate ub_colsum pass over dx); needs 256 % (c/8) == 0. */
int ub_pool2x2_bwd(const void* dy, const ub_tensor4* dcat, int c_off, const unsigned char* idx, const void* pooled,
                   void* dx, float* dbias, int n, int h, int w, int c, void* stream);

/* ---- 1x1 head, 64 -> 2 (network.py:190) and argmax (tester.py:30; trainer.py:82) ---- */
int ub_head_fwd(const void* x_bf16, const float* w, const float* bias, float* logits_nchw, unsigned char* argmax,
                int n, int h, int w_, void* stream);
/* part: partial rows [rows][194] (producer 3) = dW[2][64], db[2] and, if with_dx_sums, the column sums of dx[64] = the
 * bias gradient of conv12e */
int ub_head_bwd(const void* x_bf16, const float* w, const float* dlogits_nchw, void* dx_bf16, float* part, int with_dx_sums,
                int n, int h, int w_, void* stream);

/* ---- pixel-wise weighted loss, forward + dlogits fused (trainer.py:60-75) ----
 * kind 0: BCE-with-logits vs the one-hot target, mean over B*2*L*L (nn.BCEWithLogitsLoss(weight));
 * kind 1: weighted softmax cross-entropy, mean over B*L*L.
 * logits [B,2,hp,wp] fp32, centre-cropped by pad = (hp-L)/2; label [B,L,L] (label_bytes 1 or 8);
 * weight [B,L,L] fp32; dlogits [B,2,hp,wp] (may be NULL) receives grad_scale * dloss_sum/dz;
 * workspace: ub_loss_workspace_bytes() bytes, zero-initialised once by the caller.                  */
long long ub_loss_workspace_bytes(void);
int ub_loss_fwd_bwd(int kind, const float* logits, const void* label, int label_bytes, const float* weight,
                    float* dlogits, float* loss_out, void* workspace, int b, int hp, int wp, int l, float grad_scale,
                    void* stream);

/* ---- loss weight maps on the device (functions.py:7-78 weighted_map, functions.py:82-117 class_balance) ----
 * label: [b,h,w] uint8 (label_bytes 1) or int64 (8), binary {0,1}.  out: fp32 [b,h,w].
 *   weighted_map: w = w_c + w0*exp(-(d1+d2)^2 / (2*sigma2)) * (1 - gt), d1/d2 = exact Euclidean distance to the
 *   nearest / second-nearest 4-connected object (d2 = 0 when there is only one), w_c = count(1)/count(class);
 *   truncate_wc != 0 reproduces the reference's integer-label truncation of w_c (functions.py:37-40).
 *   The reference's constants are w0 = 20, sigma2 = 25 (functions.py:29-30).
 *   class_balance: w = count(1)/count(class) in fp32.
 * status: device int[b]; bit 0 = the sample has a single class, bit 1 = no object -- the cases in which the
 * reference raises IndexError (functions.py:40 / :60); the host wrapper turns them into that exception.
 * workspace: ub_weight_map_workspace_bytes(b,h,w) bytes of device memory (component labels + column candidates).
 * Limits: h < 2048, w <= 2048, h*w < 2^21 - 1 (packed 32-bit candidates). */
long long ub_weight_map_workspace_bytes(int b, int h, int w);
int ub_weight_map(const void* label, int label_bytes, int b, int h, int w, double w0, double sigma2, int truncate_wc,
                  float* out, void* workspace, int* status, void* stream);
int ub_class_balance(const void* label, int label_bytes, int b, int h, int w, float* out, void* workspace, int* status,
                     void* stream);

/* ---- training-sample augmentation on the device (data.py:93-137 ImageDataset.__getitem__, data.py:225-245
 * elastic_transform) for a batch of b samples ----
 * images / targets: uint8 [n_img, img_h, img_w] (targets 0 / 255).  sample: device int[b][4] = {image index, crop row
 * x, crop column y, rotation in degrees (a multiple of 30)} -- the draws of data.py:97-104,115.  noise: device fp64
 * [b][2][S][S] uniform [0,1) fields (random_state.rand(*shape) twice, data.py:236-237), S = input_size.
 * Steps: crop window, np.pad(reflect, pad_width = input_size), scipy.ndimage.rotate (cubic spline + prefilter, reshape)
 * + centre crop to S, elastic deformation (Gaussian sigma, amplitude alpha, bilinear), label crop + threshold(127),
 * min-max normalisation.  out_image: fp32 [b,1,S,S]; out_label: int64 [b,1,crop,crop] in {0,1}.
 * workspace: ub_augment_workspace_bytes(b, crop, S, sigma) bytes.  Host work: 81 Gaussian taps + 12 rotation
 * geometries, one small H2D copy; otherwise asynchronous on `stream`. */
long long ub_augment_workspace_bytes(int b, int crop, int input_size, double sigma);
int ub_augment_batch(const unsigned char* images, const unsigned char* targets, int n_img, int img_h, int img_w,
                     const int* sample, const double* noise, int b, int crop, int input_size, double alpha,
                     double sigma, float* out_image, long long* out_label, void* workspace, void* stream);

/* ---- mirror padding of test images (data.py:249-312 mirror_transform / mirror_transform_tensor) ----
 * image: fp32 [b,n,n]; out: fp32 [b,size,size], border (size-n)/2 on every side (size-n even, border < n):
 * reflected without the edge below the image, with the edge repeated above it -- the reference's slices. */
int ub_mirror_pad(const float* image, float* out, int b, int n, int size, void* stream);

/* ---- segmentation metrics on the device (functions.py:150-213: IoU, Pixel_error) ----
 * pred / label: n elements, uint8 (bytes 1) or int64 (8).  counts[3] (device, caller zeroes) receive
 * |pred & label|, |pred | label| and sum |pred - label|:  IoU = counts[0]/counts[1], pixel error = counts[2]/n. */
int ub_mask_metrics(const void* pred, int pred_bytes, const void* label, int label_bytes, long long n,
                    unsigned long long* counts, void* stream);

/* ---- overlap-tile inference plumbing (the reference only has whole-image mirror padding,
 * data.py:249-312).  Tile t = (t / nx, t % nx) produces tile_out x tile_out output pixels from a tile_in x tile_in
 * input window shifted by -halo, halo = (tile_in - tile_out) / 2.
 * ub_cut_tiles: tiles[k] (fp32 [tile_in][tile_in]) = window of tile tile_ids[k] gathered from `image`, which holds
 *   rows [row0, row0 + rows) of an h x w fp32 image; outside the image the window reflects (edge not repeated) and
 *   beyond the reflected extent repeats the last row / column.  The caller guarantees the band holds every row the
 *   listed tiles touch.  tile_ids is a device array.
 * ub_stitch_tiles: out[h][w] (uint8) receives masks[k] (uint8 [tile_out][tile_out]) at tile tile_ids[k], clipped. */
int ub_cut_tiles(const float* image, int h, int w, int row0, int rows, const int* tile_ids, int n_tiles, int nx,
                 int tile_in, int tile_out, float* tiles, void* stream);
int ub_stitch_tiles(const unsigned char* masks, const int* tile_ids, int n_tiles, int nx, int tile_out,
                    unsigned char* out, int h, int w, void* stream);

/* column sums of a bf16 [n_pix, c] view with pixel stride ld (bias gradients): partial rows [rows][c] (producer 1) */
int ub_colsum(const void* g_bf16, float* part, long long n_pix, int c, long long ld, void* stream);

/* ---- data-parallel gradient exchange over NVLink / NVSwitch peer memory (SURVEY.md 8 row e-train; csrc/allreduce.cuh) ----
 * Sum all-reduce, in place, of elements [begin, end) of an fp32 arena that every rank holds in SYMMETRIC memory:
 * peer_ptrs[k] = the arena of rank k as mapped into this process, flag_ptrs[k] = its flag pad (>= 2 * 32 * 16 uint32,
 * zeroed once), mc_ptr = the multicast address of the arenas (NVLS: in-switch reduction and broadcast through
 * multimem.ld_reduce / multimem.st) or NULL (peer loads added in rank order, one store per peer).  step_ctr / done_ctr:
 * local device uint32, zeroed once.  phases (bit mask, launched in this order on `stream`): 1 = ++step (once per
 * training step), 2 = signal "my gradients of `bucket` are final" to every peer, 4 = reduce this rank's 1/world part
 * once every rank has signalled and broadcast it, 8 = wait until all parts of buckets [0, n_buckets) have arrived.
 * Every rank must issue the same sequence; all waits are bounded. */
int ub_allreduce_sym(void* const* peer_ptrs, void* const* flag_ptrs, void* mc_ptr, unsigned int* step_ctr,
                     unsigned int* done_ctr, int rank, int world, long long begin, long long end, int bucket, int n_buckets,
                     int phases, void* stream);

/* ---- optim.SGD(lr, momentum=0.99).step() over a flat arena (trainer.py:30,78) ---- */
int ub_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, float momentum, int first_step,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H_ */
