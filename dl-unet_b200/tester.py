"""Drop-in for the reference's `tester.py`: `testing(unet, test_loader, batch_size, device,
output_dir)` writes the same images/ preds/ labels/ TIFs and test_iou.out / test_pe.out, with the
forward + centre-crop + argmax (reference tester.py:27-30) running as the fused no-grad inference
path of the B200 kernels (`Unet.predict`: argmax is computed in the head kernel's epilogue).

Multi-GPU (one process per GPU, torch.distributed initialised): the test images are independent units and
are dealt round-robin to the ranks -- no collective; every rank writes the TIFs of its own images and the
rank that owns image 0 writes the statistics (the reference's statistics are those of the first image).

`tile_in` (not in the reference, default None = the reference's whole-image forward): overlap-tile
inference -- the un-padded image is cut into tile_in x tile_in windows with a mirrored halo
(b200.tiling.TileSegmenter), the TILES are sharded over the ranks and rank 0 stitches the mask.  The
reference network zero-pads its skip connections and is therefore not tile-equivariant (SURVEY.md
0.1-2): tiled masks equal the reference run on the same tiles, not its whole-image forward.

Reference: tester.py:13-57.
"""
import os
from time import time

import numpy as np
from torchvision.utils import save_image

from functions import evaluation_metrics


def _ranks():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def testing(unet, test_loader, batch_size, device, output_dir, tile_in=None):
    start = time()
    for sub in ('images', 'preds', 'labels'):
        os.makedirs(os.path.join(output_dir, sub), exist_ok=True)
    rank, world = _ranks()
    segmenters = {}

    test_eval = None
    for idx, (image, label) in enumerate(test_loader):
        size = label.shape[-1]
        if tile_in is None:
            if idx % world != rank:
                continue                                                  # another rank's image (independent units)
            pred = unet.predict(image.to(device, non_blocking=True), size)   # [B, L, L] int64 mask
        else:
            from b200.tiling import TileSegmenter
            ipad = (image.shape[-1] - size) // 2                          # undo the loader's whole-image mirror padding
            raw = image[0, 0, ipad:ipad + size, ipad:ipad + size].float().contiguous()
            seg = segmenters.get(size)
            if seg is None:
                seg = segmenters[size] = TileSegmenter(unet, size, size, tile_in, max(1, batch_size) * 16, rank, world,
                                                       device=device)
            mask = seg(raw)                                               # stitched uint8 [L, L] on rank 0
            if rank != 0:
                continue
            pred = mask[None].long()
        pad = int((image.shape[-1] - 184 - size) / 2)                     # logits are S-184 wide (tester.py:29)
        save_image(image[0, 0, pad:size + pad, pad:size + pad], os.path.join(output_dir, 'images', f'image{idx}.tif'))
        save_image(label[0, 0, :, :].float(), os.path.join(output_dir, 'labels', f'label{idx}.tif'))
        save_image(pred[0, :, :].float(), os.path.join(output_dir, 'preds', f'pred{idx}.tif'))
        if idx == 0:
            # the reference accumulates with np.concatenate but drops the result (tester.py:42):
            # its statistics are those of the first image
            test_eval = evaluation_metrics(pred[0].detach(), label[0, 0].to(pred.device))   # reduced on the device

    if test_eval is None:                                                 # this rank does not own image 0
        return
    test = np.mean(test_eval, axis=1)
    test_std = np.std(test_eval, axis=1)
    np.savetxt(os.path.join(output_dir, 'test_iou.out'), [test[0], test_std[0]])
    np.savetxt(os.path.join(output_dir, 'test_pe.out'), [test[1], test_std[1]])

    print('Mean IoU testing:', "{:.6f}".format(test[0]))
    print('Mean PE testing :', "{:.6f}".format(test[1]))
    print('Testing took    :', "{:.6f}".format(time() - start), 's')
    print('                                                     ')
    print('Testing is finished')
