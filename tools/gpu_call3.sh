#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c3_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c3_pytest.log
timeout 600 python tools/probe2.py > $OUT/c3_probe2.log 2>&1; cp $OUT/probe2.json $OUT/c3_probe2.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c3_bench_bal.json 2> $OUT/c3_bench_bal.err
BSPB200_PANEL=1 BSPB200_GATHER=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c3_bench_bal_old.json 2> $OUT/c3_bench_bal_old.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c3_bench_grid.json 2> $OUT/c3_bench_grid.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload flat > $OUT/c3_bench_flat.json 2> $OUT/c3_bench_flat.err
tail -3 $OUT/c3_pytest.log; tail -3 $OUT/c3_probe2.log
