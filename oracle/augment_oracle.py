"""CPU oracle for the training-sample augmentation -- TEST INFRASTRUCTURE ONLY (see unet_oracle.py's header).

A numpy restatement of the reference's `ImageDataset.__getitem__` (data.py:93-137) and `elastic_transform`
(data.py:225-245) that does NOT call scipy's rotate / map_coordinates / gaussian_filter: the cubic-spline
rotation (scipy.ndimage.rotate defaults: order 3, prefilter, reshape, mode 'constant'), the separable Gaussian and
the bilinear resampling are written out, because they are the specification the device kernels follow.  The random
draws are explicit arguments (`AugDraws`), so the same draws can be replayed through the reference, this oracle
and the kernels.

Parity pin: oracle/validate_against_reference.py calls the UNMODIFIED reference `__getitem__` (numpy's global RNG
seeded, `np.random.RandomState(None)` inside elastic_transform replaced by a seeded one for the duration of the
call), records the draws it made, replays them here and requires bit-equal images and labels; the outputs are
stored in tests/golden/augment_*.npz."""
import math

import numpy as np

SQRT3M2 = math.sqrt(3.0) - 2.0            # pole of the cubic B-spline prefilter


def input_size_compute(n):
    """functions.py:121-146 for an image of size n."""
    low = 20
    size = lambda r: ((((r * 2 + 4) * 2 + 4) * 2 + 4) * 2 + 4, ((((r - 4) * 2 - 4) * 2 - 4) * 2 - 4) * 2 - 4)
    i, o = size(low)
    while o < n:
        low += 2
        i, o = size(low)
    return n, i, o


def reflect_index(i, pad, n):
    """np.pad(mode='reflect') as a gather, also when pad > n - 1 (data.py:111-112 pads by input_size)."""
    p = 2 * (n - 1)
    j = np.mod(np.asarray(i) - pad, p)
    return np.where(j >= n, p - j, j)


def cosdg_sindg(deg):
    """scipy.special.cosdg / sindg: exact at the multiples of 30 the reference draws (data.py:115)."""
    table = {0: (1.0, 0.0), 30: (math.sqrt(3) / 2, 0.5), 60: (0.5, math.sqrt(3) / 2), 90: (0.0, 1.0)}
    q, r = divmod(int(deg) % 360, 90)
    c, s = table[r]
    for _ in range(q):                    # rotate by 90 degrees: (c, s) -> (-s, c)
        c, s = -s, c
    return c, s


def spline_prefilter(a):
    """scipy.ndimage.spline_filter(order=3, mode='constant' -> mirror boundary), both axes, float64."""
    a = np.array(a, dtype=np.float64)
    z = SQRT3M2
    for axis in (0, 1):
        a = np.moveaxis(a, axis, 0)
        n = a.shape[0]
        a *= (1 - z) * (1 - 1 / z)
        horizon = min(n, int(math.ceil(math.log(1e-17) / math.log(abs(z)))))
        a[0] = np.tensordot(z ** np.arange(horizon), a[:horizon], axes=(0, 0))
        for k in range(1, n):
            a[k] += z * a[k - 1]
        a[n - 1] = (z / (z * z - 1)) * (a[n - 1] + z * a[n - 2])
        for k in range(n - 2, -1, -1):
            a[k] = z * (a[k + 1] - a[k])
        a = np.moveaxis(a, 0, axis)
    return a


def bspline3(t):
    return ((1 - t) ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6, (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6, t ** 3 / 6)


def to_u8(v):
    """scipy's float -> uint8 output cast: round half up, clip."""
    return np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)


def rotation_geometry(in_size, deg, out_size):
    """scipy.ndimage.rotate(reshape=True) geometry + the centre crop of data.py:118-124.  Returns (R, offset, t, l):
    output pixel (i, j) of the crop reads the input at R @ (t + i, l + j) + offset."""
    c, s = cosdg_sindg(deg)
    R = np.array([[c, s], [-s, c]])
    ob = R @ np.array([[0, 0, in_size, in_size], [0, in_size, 0, in_size]], dtype=np.float64)
    oshape = (np.ptp(ob, axis=1) + 0.5).astype(int)
    off = (np.array([in_size, in_size]) - 1) / 2 - R @ ((oshape - 1) / 2)
    h, w = int(oshape[0]), int(oshape[1])
    return R, off, h // 2 - out_size // 2, w // 2 - out_size // 2


def rotate_center_crop(img_u8, deg, out_size):
    """rotate(img, deg)[t:b, l:r] (data.py:116-124) for a uint8 image."""
    n = img_u8.shape[0]
    R, off, t, l = rotation_geometry(n, deg, out_size)
    coef = spline_prefilter(img_u8)
    oy, ox = np.meshgrid(np.arange(t, t + out_size), np.arange(l, l + out_size), indexing="ij")
    yy = R[0, 0] * oy + R[0, 1] * ox + off[0]
    xx = R[1, 0] * oy + R[1, 1] * ox + off[1]
    y0, x0 = np.floor(yy).astype(int), np.floor(xx).astype(int)
    wy, wx = bspline3(yy - y0), bspline3(xx - x0)
    out = np.zeros(yy.shape)
    for a in range(4):
        for b in range(4):
            out += wy[a] * wx[b] * coef[np.clip(y0 - 1 + a, 0, n - 1), np.clip(x0 - 1 + b, 0, n - 1)]
    return to_u8(out)


def gaussian_kernel(sigma, truncate=4.0):
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1)
    k = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return k / k.sum()


def gaussian_filter_zero(a, sigma):
    """scipy.ndimage.gaussian_filter(a, sigma, mode='constant', cval=0): axis 0, then axis 1 (data.py:236-237)."""
    k = gaussian_kernel(sigma)
    r = len(k) // 2
    for axis in (0, 1):
        a = np.moveaxis(a, axis, 0)
        p = np.concatenate([np.zeros((r,) + a.shape[1:]), a, np.zeros((r,) + a.shape[1:])])
        out = np.zeros_like(a, dtype=np.float64)
        # scipy's correlate1d on a symmetric kernel sums the outer taps pairwise around the centre
        for j in range(1, r + 1):
            out += k[r + j] * (p[r - j:r - j + a.shape[0]] + p[r + j:r + j + a.shape[0]])
        out += k[r] * p[r:r + a.shape[0]]
        a = np.moveaxis(out, 0, axis)
    return a


def bilinear_constant(img_u8, yy, xx):
    """map_coordinates(img, (yy, xx), order=1, mode='constant', cval=0) for a uint8 image (data.py:245): a
    coordinate outside [0, n-1] gives cval."""
    n, m = img_u8.shape
    inside = (yy >= 0) & (yy <= n - 1) & (xx >= 0) & (xx <= m - 1)
    yc, xc = np.where(inside, yy, 0.0), np.where(inside, xx, 0.0)
    y0, x0 = np.floor(yc).astype(int), np.floor(xc).astype(int)
    ty, tx = yc - y0, xc - x0
    y1, x1 = np.minimum(y0 + 1, n - 1), np.minimum(x0 + 1, m - 1)
    f = img_u8.astype(np.float64)
    # scipy accumulates the 4 taps in the order (0,0), (0,1), (1,0), (1,1) with weights wy * wx
    v = (1 - ty) * (1 - tx) * f[y0, x0]
    v = v + (1 - ty) * tx * f[y0, x1]
    v = v + ty * (1 - tx) * f[y1, x0]
    v = v + ty * tx * f[y1, x1]
    return to_u8(np.where(inside, v, 0.0))


def elastic_transform(images, alpha, sigma, noise_x, noise_y):
    """data.py:225-245 with the two uniform [0,1) fields passed in (random_state.rand(*shape), twice)."""
    shape = images[0].shape
    dx = gaussian_filter_zero(noise_x * 2 - 1, sigma) * alpha
    dy = gaussian_filter_zero(noise_y * 2 - 1, sigma) * alpha
    x, y = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    return [bilinear_constant(im, x + dx, y + dy) for im in images]


def augment_sample(image, target, x, y, rot_deg, noise_x, noise_y, crop=388, alpha=3, sigma=10):
    """data.py:93-137 after the random draws: crop window at (x, y), reflect pad by input_size, rotate + centre
    crop, elastic deformation, label crop + threshold, min-max normalisation.
    -> (inp float32 [input_size, input_size], gt int64 [crop, crop])."""
    img = image[x:x + crop, y:y + crop]
    tgt = target[x:x + crop, y:y + crop]
    original, input_size, _ = input_size_compute(img.shape[-1])
    idx = reflect_index(np.arange(original + 2 * input_size), input_size, original)
    img_pad, tgt_pad = img[idx][:, idx], tgt[idx][:, idx]                      # data.py:111-112
    img_rot = rotate_center_crop(img_pad, rot_deg, input_size)                # data.py:115-124
    tgt_rot = rotate_center_crop(tgt_pad, rot_deg, input_size)
    inp, gt = elastic_transform((img_rot, tgt_rot), alpha, sigma, noise_x, noise_y)   # data.py:127
    pad = int((input_size - original) / 2)
    gt = gt[pad:original + pad, pad:original + pad]                           # data.py:131
    gt = (gt > 127).astype(np.int64)                                          # threshold(127) / 255 -> long
    inp = inp.astype(np.float64)
    inp = (inp - inp.min()) / np.ptp(inp)                                     # data.py:134
    return inp.astype(np.float32), gt


def synthetic_dataset_image(size, seed):
    """A seeded stand-in for one dataset frame: smooth uint8 texture + 0/255 blob mask (data.py:63's gt_bin)."""
    from scipy import ndimage
    rng = np.random.RandomState(seed)
    img = ndimage.gaussian_filter(rng.rand(size, size), 3)
    img = ((img - img.min()) / np.ptp(img) * 255).astype(np.uint8)
    yy, xx = np.indices((size, size))
    gt = np.zeros((size, size), np.uint8)
    for _ in range(25):
        cy, cx = rng.rand(2) * size
        ry, rx = size / 64 + rng.rand(2) * size / 12
        gt[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = 255
    return img, gt


def noise_fields(seed, size):
    """The two random_state.rand(*shape) fields of data.py:236-237 for RandomState(seed)."""
    rs = np.random.RandomState(seed)
    return np.stack([rs.rand(size, size), rs.rand(size, size)])
