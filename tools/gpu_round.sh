#!/bin/bash
# One GPU-box pass: tests, bench lines of every workload, ncu launch list / DRAM single pass / --set full.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
tag=${1:-r02}
mkdir -p gpurun_out
O=gpurun_out
if [ "${2:-tests}" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log
  tail -3 $O/${tag}_pytest.log
fi
timeout 400 python bench.py > $O/${tag}_bench_cfg3.json 2> $O/${tag}_bench_cfg3.err; tail -c 600 $O/${tag}_bench_cfg3.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_ref.json 2> $O/${tag}_bench_ref.err
for w in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --no-cpu > $O/${tag}_bench_$w.json 2> $O/${tag}_bench_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $O/${tag}_launches.csv \
  python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-single --no-formats > $O/${tag}_ncu1.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
  -s 150 -c 200 --csv --log-file $O/${tag}_dram.csv \
  python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-single --no-formats > $O/${tag}_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_inspectors|k_chan_ifft16" -s 4 -c 2 -f -o $O/${tag}_full \
  python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu --no-single --no-formats > $O/${tag}_ncu3.log 2>&1
ls -la $O | tail -20
