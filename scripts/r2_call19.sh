#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/c19_pytest_k.log 2>&1; tail -5 gpurun_out/c19_pytest_k.log | cut -c1-250
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_kernels.py > gpurun_out/c19_pytest.log 2>&1; tail -5 gpurun_out/c19_pytest.log | cut -c1-250
timeout 600 python tests/prof_layers.py 16 > gpurun_out/c19_layers.log 2>&1; tail -26 gpurun_out/c19_layers.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c19_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['inference']['value'], d['inference']['e2e']['value'])
r=d['roofline']; print('igemm', r['frac'], r['ms_per_step'], 'wgrad', r['wgrad_kernel']['frac'], r['wgrad_kernel']['ms_per_step'])
PY
tail -3 gpurun_out/c19_bench.err
UB_OPTIONS="11=0" timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/c19_bench_nopair.json 2> gpurun_out/c19_bench_nopair.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c19_bench_nopair.json'))
r=d['roofline']
print('no pair:', d['value'], d['ms_per_step'], 'igemm', r['ms_per_step'], 'wgrad', r['wgrad_kernel']['ms_per_step'])
PY
