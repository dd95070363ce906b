#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c3_pytest.log 2>&1
tail -15 gpurun_out/c3_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
cat gpurun_out/c3_bench.json; tail -5 gpurun_out/c3_bench.err
