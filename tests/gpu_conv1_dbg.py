"""bring-up bisect of the tensor-pipe conv11c wgrad: python tests/gpu_conv1_dbg.py <dbg bits>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch
from b200 import ops
dbg = int(sys.argv[1])
n, h, w = 2, 200, 264
ops.set_option(9, dbg)
img = torch.rand(n, 1, h, w, device="cuda")
dy = torch.randn(n, h - 2, w - 2, 64, device="cuda").to(torch.bfloat16)
rows = ops.partial_rows(ops.PROD_CONV1, n, h, w)
part = torch.zeros(rows, 640, device="cuda")
ops.conv1_wgrad(img, dy, part=part)
torch.cuda.synchronize()
rc, code = ops.device_error()
print("dbg", dbg, "ok rows", rows, "sum", float(part.sum()), "device_error", rc, hex(code))
