"""GPU parity of the device-side augmentation (csrc/augment.cuh through ub_augment_batch)
against oracle/augment_oracle.py and the golden outputs of the reference's own ImageDataset.__getitem__ (data.py:93-137).

The pipeline ends in uint8 roundings (scipy's output casts), so the device result equals the oracle exactly except
where an fp64 value lands within ~1e-13 of a rounding boundary: allowed are differences of one grey level on < 1e-4 of
the pixels and label flips on < 1e-4 of the label pixels; everything else must be identical."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def compare(img, lab, o_img, o_lab):
    img, lab = img.cpu().numpy(), lab.cpu().numpy()
    assert img.dtype == np.float32 and lab.dtype == np.int64
    # one grey level of the normalised image = 1 / ptp >= 1/255
    d = np.abs(img.astype(np.float64) - o_img.astype(np.float64))
    assert d.max() <= 1.0 / 255 * 1.51 + 1e-6, d.max()          # min/max could each move by a level in the worst case
    assert (img != o_img).mean() < 1e-4, (img != o_img).mean()
    assert (lab != o_lab).mean() < 1e-4, (lab != o_lab).mean()


@pytest.mark.parametrize("name", ["augment_crop100", "augment_crop388"])
def test_augment_matches_reference_golden(golden_dir, name):
    import augment_oracle as A
    from b200.augment import DeviceAugmenter
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    size, crop, seed = int(g["size"]), int(g["crop"]), int(g["seed"])
    img, gt = A.synthetic_dataset_image(size, seed)
    aug = DeviceAugmenter(img[None], gt[None], crop=crop)
    S = aug.input_size
    noise = A.noise_fields(int(g["noise_seed"]), S)
    out_img, out_lab = aug(draws=g["draws"][None], noise=noise[None])
    assert out_img.shape == (1, 1, S, S) and out_lab.shape == (1, 1, crop, crop)
    compare(out_img[0, 0], out_lab[0, 0], g["image"], g["label"].astype(np.int64))


def test_augment_batch_all_rotations_vs_oracle():
    """A batch of 12 samples, one per rotation angle, from two dataset frames, different crop windows."""
    import augment_oracle as A
    from b200.augment import DeviceAugmenter
    crop, size = 100, 180
    frames = [A.synthetic_dataset_image(size, s) for s in (11, 12)]
    aug = DeviceAugmenter(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), crop=crop)
    S = aug.input_size
    rng = np.random.RandomState(5)
    draws = np.array([[k % 2, rng.randint(0, size - crop + 1), rng.randint(0, size - crop + 1), 30 * k] for k in range(12)])
    noise = np.stack([A.noise_fields(100 + k, S) for k in range(12)])
    out_img, out_lab = aug(draws=draws, noise=noise)
    for k in range(12):
        o_img, o_lab = A.augment_sample(frames[draws[k, 0]][0], frames[draws[k, 0]][1], draws[k, 1], draws[k, 2], draws[k, 3],
                                        noise[k, 0], noise[k, 1], crop=crop)
        compare(out_img[k, 0], out_lab[k, 0], o_img, o_lab)


def test_augment_draws_follow_reference_order_and_feed_training():
    """draw() consumes numpy's global RNG exactly like data.py:97-104,115; the default path (device noise) yields
    tiles the network trains on."""
    import augment_oracle as A
    from b200.augment import DeviceAugmenter
    from network import Unet
    from b200 import train as T
    img, gt = A.synthetic_dataset_image(300, 21)
    aug = DeviceAugmenter(img[None], gt[None], crop=36)          # input_size_compute(36): 380 in, 196 out >= 36
    np.random.seed(9)
    d = aug.draw([0, 0, 0])
    np.random.seed(9)
    for k in range(3):
        crop_id = np.random.choice(range(len(aug.pairs)), 1, p=aug.distribution[0])[0]
        x, y = aug.pairs[crop_id]
        x += np.random.randint(-5, 6)
        y += np.random.randint(-5, 6)
        rot = np.random.choice(np.arange(0, 360, 30))
        assert tuple(d[k]) == (0, min(max(0, x), 300 - 36), min(max(0, y), 300 - 36), rot)
    images, labels = aug([0, 0])
    assert aug.input_size == 380 and images.shape == (2, 1, 380, 380) and labels.shape == (2, 1, 36, 36)
    assert float(images.min()) == 0.0 and float(images.max()) == 1.0 and set(labels.unique().tolist()) <= {0, 1}
    torch.manual_seed(0)
    net = Unet().to(DEV)
    step = T.TrainStep(net, T.FusedSGD(net), graph=False)
    w = torch.ones(2, 36, 36, device=DEV)
    assert torch.isfinite(step(images, labels, w))
    with pytest.raises(RuntimeError):
        aug(draws=np.array([[0, 0, 0, 45]]))                     # not one of the reference's angles
