#!/bin/bash
mkdir -p gpurun_out
for d in 0 16; do
  timeout 60 python tests/gpu_conv1_dbg.py $d 2>&1 | tail -2 | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv1 or unpack" 2>&1 | tail -5
