#!/bin/bash
set -x
mkdir -p gpurun_out
python tests/gpu_b16_parity.py 16 > gpurun_out/c4_b16_parity.log 2>&1; cat gpurun_out/c4_b16_parity.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect "tests/test_gpu_network.py::test_bench_config_parity" > gpurun_out/c4_pytest.log 2>&1
tail -15 gpurun_out/c4_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
cat gpurun_out/c4_bench.json; tail -5 gpurun_out/c4_bench.err
