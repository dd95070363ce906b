#!/bin/bash
# eight GPUs: the N = 8 and N = 4 bench lines with the final kernels (the launch the driver uses)
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/c49_bench_n$n.json 2> gpurun_out/c49_bench_n$n.err
echo "bench n$n rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/c49_bench_n$n.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"])
sb=d.get("scaling_breakdown",{})
print({k:sb.get(k) for k in ("compute_only_eager_ms","compute_only_graph_ms","exposed_allreduce_ms","n1_equivalent_tiles_per_s")})
i=d.get("inference",{})
print(i.get("value"), i.get("e2e",{}).get("value"))
PY
done
