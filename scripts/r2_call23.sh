#!/bin/bash
# fused column sums in the CTA-pair dgrads: kernel tests, network tests, bench A/B is against call 20 (13.92 ms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dgrad or upconv or pair" > gpurun_out/c23_kernels.log 2>&1
echo "kernels rc=$?"; tail -5 gpurun_out/c23_kernels.log
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/c23_net.log 2>&1
echo "net rc=$?"; tail -5 gpurun_out/c23_net.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/c23_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
for k,v in d["roofline"]["families"].items(): print(k, v)
PY
