#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c5_pytest.log 2>&1
tail -15 gpurun_out/c5_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
cat gpurun_out/c5_bench.json | head -c 1500; tail -5 gpurun_out/c5_bench.err
UB_UNPACK_PER_LAYER=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/c5_bench_perlayer.json 2> gpurun_out/c5_bench_perlayer.err
cat gpurun_out/c5_bench_perlayer.json | head -c 600; tail -5 gpurun_out/c5_bench_perlayer.err
