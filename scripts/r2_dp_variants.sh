#!/bin/bash
# data-parallel schedule variants at N GPUs (default 2), 16 tiles per GPU (the per-GPU load of the 8-GPU configuration)
N=${1:-2}
B=${2:-16}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --batch-per-gpu $B --no-inference > gpurun_out/dp_${N}_${name}.json 2> gpurun_out/dp_${N}_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp_${N}_${name}.json"))
    print("${name}", round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms", d.get("scaling_breakdown"))
except Exception as e:
    print("${name} failed", e)
PY
  tail -3 gpurun_out/dp_${N}_${name}.err
}
run base UB_X=0
run nccl_info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL UB_X=0
grep -E "NVLS|channels|Channel|Ring|Tree|nChannels" gpurun_out/dp_${N}_nccl_info.err | head -20
run reserve16 UB_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16
run reserve8 UB_DP_RESERVE_SMS=8 NCCL_MAX_CTAS=8 UB_DP_RESERVE_KERNELS=3
run reserve32 UB_DP_RESERVE_SMS=32 NCCL_MAX_CTAS=32 UB_DP_RESERVE_KERNELS=2
run maxctas8 NCCL_MAX_CTAS=8
run graph UB_DP_GRAPH=1
run graph_reserve16 UB_DP_GRAPH=1 UB_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16
