#!/bin/bash
# cfg5 check: the tests that touch the sweep / detector / view, the bench line, the launch list
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_panoramic.py tests/test_gpu_view_snr.py tests/test_gpu_chdet.py tests/test_golden.py tests/test_gpu_analyzer.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --workload cfg5 --no-cpu > $O/c5.json 2> $O/c5.err
python -c "
import json; d=json.load(open('$O/c5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['phases_ms'])"; tail -3 $O/c5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/c5_launches.csv python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu > $O/c5_ncu.log 2>&1
