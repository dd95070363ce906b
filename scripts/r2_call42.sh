#!/bin/bash
# running column sums + HSET2 mask in the 16-warp dgrad epilogue, fused pool as a two-stage (drain / pool) epilogue:
# parity, then A/B against the previous commit's library (libunet_b200_base.so) on the same box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c42_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/c42_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sizes.py -x -q -m gpu > gpurun_out/c42_net.log 2>&1
echo "net rc=$?"; tail -3 gpurun_out/c42_net.log
for i in 1 2; do
for v in new base; do
if [ $v = base ]; then export UB_LIB_PATH=$PWD/dl-unet_b200/csrc/libunet_b200_base.so; else unset UB_LIB_PATH; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-gpu-baseline > gpurun_out/c42_bench_${v}_$i.json 2> gpurun_out/c42_bench_${v}_$i.err
echo "bench $v rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c42_bench_${v}_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
f=d["roofline"]["families"]
print({k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("igemm")})
PY
done
done
unset UB_LIB_PATH
timeout 600 python tests/prof_epi_clocks.py 16 > gpurun_out/c42_clocks.log 2>&1
echo "prof rc=$?"; grep "pairs 1" gpurun_out/c42_clocks.log | cut -c1-40,100-400
