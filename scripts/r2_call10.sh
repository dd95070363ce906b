#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv1 or unpack" > gpurun_out/c10_pytest_k.log 2>&1; tail -5 gpurun_out/c10_pytest_k.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py -x -q > gpurun_out/c10_pytest_n.log 2>&1; tail -5 gpurun_out/c10_pytest_n.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c10_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in d['roofline']['families'].items():
    if k.startswith('hbm'): print(k, round(v['ms_per_step'],3), round(v.get('hbm_frac',0),3))
PY
tail -3 gpurun_out/c10_bench.err
