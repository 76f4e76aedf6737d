#!/bin/bash
# 8-GPU check of the two scaling lines (cfg3 weak, cfg5 strong)
mkdir -p gpurun_out; O=gpurun_out
N=${1:-8}
for w in cfg3 cfg5; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --workload $w --no-cpu --no-formats --no-single > $O/n${N}_$w.json 2> $O/n${N}_$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/n${N}_$w.json") if l.startswith("{")][-1])
    print("$w", d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d.get("host_binding"), d["roofline"].get("phases_ms"), d["roofline"].get("gather_bytes"))
except Exception as ex: print("$w failed", ex)
PY
  tail -2 $O/n${N}_$w.err | cut -c1-300
done
