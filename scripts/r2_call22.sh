#!/bin/bash
N=4
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-inference > gpurun_out/c22_n4_${name}.json 2> gpurun_out/c22_n4_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c22_n4_${name}.json"))
    b=d.get("scaling_breakdown") or {}
    print("${name}", round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("step_ms_min_median_max"))
except Exception as e:
    print("${name} failed", e)
PY
}
run symm UB_DP_ALLREDUCE=symm
run peer UB_DP_ALLREDUCE=peer
run nccl UB_DP_ALLREDUCE=nccl
run symm2 UB_DP_ALLREDUCE=symm
