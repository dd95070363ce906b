#!/bin/bash
N=${1:-2}
B=${2:-16}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --batch-per-gpu $B --no-inference > gpurun_out/dp2_${N}_${name}.json 2> gpurun_out/dp2_${N}_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp2_${N}_${name}.json"))
    b=d.get("scaling_breakdown") or {}
    print("${name}", round(d["value"],1), "tiles/s", round(d["ms_per_step"],3), "ms | eager", round(b.get("compute_only_eager_ms",0),3), "graph", round(b.get("compute_only_graph_ms",0),3), "exposed", round(b.get("exposed_allreduce_ms",0),3), "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("${name} failed", e)
PY
  tail -2 gpurun_out/dp2_${N}_${name}.err | cut -c1-300
}
run b32 UB_DP_BUCKET_MB=32
run b64 UB_DP_BUCKET_MB=64
run b1e9 UB_DP_BUCKET_MB=1000000
run b32_graph UB_DP_BUCKET_MB=32 UB_DP_GRAPH=1
run b1e9_graph UB_DP_BUCKET_MB=1000000 UB_DP_GRAPH=1
run b16_graph UB_DP_BUCKET_MB=16 UB_DP_GRAPH=1
