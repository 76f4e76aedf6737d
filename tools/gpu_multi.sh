#!/bin/bash
# two-GPU box: the NCCL tests + cfg5 / cfg3 bench lines at N = 2
mkdir -p gpurun_out; O=gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_panoramic.py -m gpu -x -q 2>&1 | tail -4
for w in cfg5 cfg3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $w --no-cpu --no-formats --no-single > $O/multi_${w}_n$N.json 2> $O/multi_${w}_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/multi_${w}_n$N.json") if l.startswith("{")][-1])
    print("$w", d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"].get("phases_ms"), d["roofline"].get("gather_bytes"))
except Exception as ex: print("$w failed", ex)
PY
  tail -2 $O/multi_${w}_n$N.err
done
