#!/bin/bash
# CTA-pair kernel, forward flavour: two chunks per epilogue pass -- parity, A/B against the previous commit (libunet_b200_base.so), up-conv timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c51_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c51_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py -x -q -m gpu > gpurun_out/c51_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c51_net.log
for i in 1 2; do
for v in new base; do
if [ $v = base ]; then export UB_LIB_PATH=$PWD/dl-unet_b200/csrc/libunet_b200_base.so; else unset UB_LIB_PATH; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c51_bench_${v}_$i.json 2> gpurun_out/c51_bench_${v}_$i.err
echo "bench $v rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c51_bench_${v}_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
f=d["roofline"]["families"]
print({k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("igemm")})
PY
done
done
unset UB_LIB_PATH
for l in 1 2 3 4; do for v in new base; do
if [ $v = base ]; then export UB_LIB_PATH=$PWD/dl-unet_b200/csrc/libunet_b200_base.so; else unset UB_LIB_PATH; fi
echo "level $l $v"; timeout 300 python tests/prof_upconv.py $l 16 | head -2; done; done > gpurun_out/c51_upconv.log 2>&1
unset UB_LIB_PATH; cat gpurun_out/c51_upconv.log
