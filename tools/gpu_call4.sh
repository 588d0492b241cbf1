#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 tools/ubench/fp64_ubench > $OUT/c4_ubench.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c4_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c4_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c4_bench_bal.json 2> $OUT/c4_bench_bal.err
BSPB200_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c4_bench_bal_nopdl.json 2> $OUT/c4_bench_bal_nopdl.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c4_bench_grid.json 2> $OUT/c4_bench_grid.err
cat $OUT/c4_ubench.log; tail -3 $OUT/c4_pytest.log
