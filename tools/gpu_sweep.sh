#!/bin/bash
# stream-count sweep of cfg3 + role counters of the instrumented twin
mkdir -p gpurun_out; O=gpurun_out
for s in 148 222 296 444 592; do
  timeout 200 python bench.py --workload cfg3 --streams $s --no-cpu --no-formats --no-single > $O/sweep_s$s.json 2> $O/sweep_s$s.err
  python - <<PY
import json
try:
    d=json.load(open("$O/sweep_s$s.json")); r=d["roofline"]
    print($s, round(d["value"]), d["ms_per_step"], r["device_ms_per_step"], round(d["e2e"]["value"]))
except Exception as ex: print($s, "failed", ex)
PY
done
SDB_LIB=sigdigger_b200/libsigdigger_b200_prof.so timeout 200 python bench.py --workload cfg3 --no-cpu --no-formats --no-single > $O/prof_cfg3.json 2> $O/prof_cfg3.err
python - <<PY
import json
d=json.load(open("$O/prof_cfg3.json")); r=d["roofline"]
print(r.get("inspector_role_cycles_per_sample")); print(r.get("inspector_cta_by_class"))
PY
