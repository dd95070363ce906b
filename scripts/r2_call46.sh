#!/bin/bash
# image-spanning boxes also for the up-convolutions: parity, then option 18 on / off on the same box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c46_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c46_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py -x -q -m gpu > gpurun_out/c46_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c46_net.log
for i in 1; do
for opt in "18=1" "18=0"; do
UB_OPTIONS=$opt timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c46_bench_${opt/=/_}_$i.json 2> gpurun_out/c46_bench_${opt/=/_}_$i.err
echo "bench $opt rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c46_bench_${opt/=/_}_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
f=d["roofline"]["families"]
print({k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("igemm")})
PY
done
done
for l in 3 4; do for o in 1 0; do echo "level $l option18=$o"; UB_OPTIONS="18=$o" timeout 300 python tests/prof_upconv.py $l 16 | tail -4; done; done > gpurun_out/c46_upconv.log 2>&1; cat gpurun_out/c46_upconv.log
