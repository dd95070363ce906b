#!/bin/bash
# one-CTA igemm kernel: fused column sums for up-convolution 1's dgrad (the last colsum_kernel launch of the step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c54_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c54_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/c54_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c54_net.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c54_bench_$i.json 2> gpurun_out/c54_bench_$i.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c54_bench_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
f=d["roofline"]["families"]
print({k:round(v["ms_per_step"],3) for k,v in f.items() if "up" in k or "colsum" in k})
PY
done
