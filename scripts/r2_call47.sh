#!/bin/bash
# wgrad4: nine taps in ten slots (option 19) -- parity, then A/B on the same box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c47_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c47_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py -x -q -m gpu > gpurun_out/c47_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c47_net.log
for i in 1 2; do
for opt in "19=1" "19=0"; do
UB_OPTIONS=$opt timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c47_bench_${opt/=/_}_$i.json 2> gpurun_out/c47_bench_${opt/=/_}_$i.err
echo "bench $opt rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c47_bench_${opt/=/_}_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
f=d["roofline"]["families"]
print({k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("wgrad")})
PY
done
done
