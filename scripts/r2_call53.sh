#!/bin/bash
# two GPUs: data-parallel gradient equality tests and the N = 2 bench with the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu -k "data_parallel" > gpurun_out/c53_dp_tests.log 2>&1
echo "dp tests rc=$?"; tail -3 gpurun_out/c53_dp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c53_bench_n2.json 2> gpurun_out/c53_bench_n2.err
echo "bench n2 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c53_bench_n2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d.get("scaling_breakdown"), d.get("inference"))
PY
