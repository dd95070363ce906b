#!/bin/bash
# round-2: persistent deterministic wgrad -- parity tests, layer table, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "wgrad or upconv or unpack" > gpurun_out/c2_pytest_kernels.log 2>&1
tail -5 gpurun_out/c2_pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_dropin.py -x -q > gpurun_out/c2_pytest_net.log 2>&1
tail -5 gpurun_out/c2_pytest_net.log
timeout 600 python tests/prof_layers.py 16 > gpurun_out/c2_layers.log 2>&1
tail -30 gpurun_out/c2_layers.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
cat gpurun_out/c2_bench.json
