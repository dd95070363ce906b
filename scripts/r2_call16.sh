#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv1" > gpurun_out/c16_pytest_k.log 2>&1; tail -12 gpurun_out/c16_pytest_k.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k "not data_parallel" > gpurun_out/c16_pytest_n.log 2>&1; tail -4 gpurun_out/c16_pytest_n.log | cut -c1-250
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c16_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['inference']['value'], d['inference']['e2e']['value'])
for k,v in d['roofline']['families'].items():
    if k.startswith('hbm'): print(k, round(v['ms_per_step'],3), round(v.get('hbm_frac',0),3))
PY
tail -3 gpurun_out/c16_bench.err
