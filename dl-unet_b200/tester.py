"""Drop-in for the reference's `tester.py`: `testing(unet, test_loader, batch_size, device,
output_dir)` writes the same images/ preds/ labels/ TIFs and test_iou.out / test_pe.out, with the
forward + centre-crop + argmax (reference tester.py:27-30) running as the fused no-grad inference
path of the B200 kernels (`Unet.predict`: argmax is computed in the head kernel's epilogue).

Reference: tester.py:13-57.
"""
import os
from time import time

import numpy as np
from torchvision.utils import save_image

from functions import evaluation_metrics


def testing(unet, test_loader, batch_size, device, output_dir):
    start = time()
    for sub in ('images', 'preds', 'labels'):
        os.makedirs(os.path.join(output_dir, sub), exist_ok=True)

    test_eval = None
    for idx, (image, label) in enumerate(test_loader):
        size = label.shape[-1]
        pred = unet.predict(image.to(device, non_blocking=True), size)   # [B, L, L] int64 mask
        pad = int((image.shape[-1] - 184 - size) / 2)                     # logits are S-184 wide (tester.py:29)
        save_image(image[0, 0, pad:size + pad, pad:size + pad], os.path.join(output_dir, 'images', f'image{idx}.tif'))
        save_image(label[0, 0, :, :].float(), os.path.join(output_dir, 'labels', f'label{idx}.tif'))
        save_image(pred[0, :, :].float(), os.path.join(output_dir, 'preds', f'pred{idx}.tif'))
        if test_eval is None:
            # the reference accumulates with np.concatenate but drops the result (tester.py:42):
            # its statistics are those of the first image
            test_eval = evaluation_metrics(pred[0].detach(), label[0, 0].to(pred.device))   # reduced on the device

    test = np.mean(test_eval, axis=1)
    test_std = np.std(test_eval, axis=1)
    np.savetxt(os.path.join(output_dir, 'test_iou.out'), [test[0], test_std[0]])
    np.savetxt(os.path.join(output_dir, 'test_pe.out'), [test[1], test_std[1]])

    print('Mean IoU testing:', "{:.6f}".format(test[0]))
    print('Mean PE testing :', "{:.6f}".format(test[1]))
    print('Testing took    :', "{:.6f}".format(time() - start), 's')
    print('                                                     ')
    print('Testing is finished')
