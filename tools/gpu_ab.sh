#!/bin/bash
# A/B of the round's last kernel changes on cfg3, then the whole GPU suite
mkdir -p gpurun_out; O=gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --workload cfg3 --no-cpu --no-formats --no-single > $O/ab_$tag.json 2> $O/ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/ab_$tag.json")); r=d["roofline"]
    print("$tag", round(d["value"]), round(d["ms_per_step"],3), r["device_ms_per_step"])
except Exception as ex: print("$tag failed", ex)
PY
}
run base SDB_IFFT16_OCC=3
run occ4 SDB_IFFT16_OCC=4
run base2 SDB_IFFT16_OCC=3
run occ4b SDB_IFFT16_OCC=4
SDB_LIB=sigdigger_b200/libsigdigger_b200_prof.so timeout 200 python bench.py --workload cfg3 --no-cpu --no-formats --no-single > $O/ab_prof.json 2> $O/ab_prof.err
python -c "
import json; r=json.load(open('$O/ab_prof.json'))['roofline']; print(r.get('inspector_role_cycles_per_sample')); print({k:(round(v['mean_cycles']),v['max_cycles'],v['role_cycles_per_sample']) for k,v in r.get('inspector_cta_by_class',{}).items()})"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
