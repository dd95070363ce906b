"""Device augmentation timing (and the reference-equivalent CPU route beside it): python tests/prof_augment.py [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from b200.augment import DeviceAugmenter  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.RandomState(0)
imgs = (rng.rand(4, 512, 512) * 255).astype(np.uint8)
yy, xx = np.indices((512, 512))
gts = np.zeros((4, 512, 512), np.uint8)
for k in range(4):
    for _ in range(25):
        cy, cx = rng.rand(2) * 512
        ry, rx = 8 + rng.rand(2) * 40
        gts[k][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = 255
aug = DeviceAugmenter(imgs, gts, crop=388)
idx = [i % 4 for i in range(B)]
for _ in range(2):
    aug(idx)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    x, y = aug(idx)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"DeviceAugmenter batch {B} x 388-crop -> 572 tiles: {dt * 1e3:.2f} ms/batch = {B / dt:.0f} tiles/s "
      f"(host draws + device noise + 10 kernels)")
if "--cpu" in sys.argv:
    # the reference's route for ONE sample: scipy rotate of the 1532^2 padded image (x2), gaussian_filter, map_coordinates
    from scipy.ndimage import gaussian_filter, map_coordinates, rotate
    t0 = time.perf_counter()
    pad_i, pad_t = np.pad(imgs[0][:388, :388], 572, mode="reflect"), np.pad(gts[0][:388, :388], 572, mode="reflect")
    ri, rt = rotate(pad_i, 30), rotate(pad_t, 30)
    h, w = ri.shape
    ri, rt = ri[h // 2 - 286:h // 2 + 286, w // 2 - 286:w // 2 + 286], rt[h // 2 - 286:h // 2 + 286, w // 2 - 286:w // 2 + 286]
    rs = np.random.RandomState(0)
    dx = gaussian_filter(rs.rand(572, 572) * 2 - 1, 10, mode="constant", cval=0) * 3
    dy = gaussian_filter(rs.rand(572, 572) * 2 - 1, 10, mode="constant", cval=0) * 3
    gx, gy = np.meshgrid(np.arange(572), np.arange(572), indexing="ij")
    ind = np.reshape(gx + dx, (-1, 1)), np.reshape(gy + dy, (-1, 1))
    a, b = map_coordinates(ri, ind, order=1), map_coordinates(rt, ind, order=1)
    cpu = time.perf_counter() - t0
    print(f"CPU scipy route (data.py:108-127 for one sample): {cpu * 1e3:.0f} ms/sample = {1 / cpu:.2f} tiles/s per core; "
          f"device / CPU-core = {B / dt * cpu:.0f}x")
