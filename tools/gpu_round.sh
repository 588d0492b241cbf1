#!/bin/bash
# One GPU call: parity tests, bench on every workload, ncu launch list of the headline bench. Output -> gpurun_out/
set -u
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_bal.json 2> $OUT/${TAG}_bench_bal.err
for wl in flat bal_small stress grid; do
  timeout 400 python bench.py --steps 5 --warmup 3 --workload $wl > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
head -c 600 $OUT/${TAG}_bench_bal.json
