#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c10_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c10_pytest.log
for lanes in 4 1 8; do
BSPB200_LANES=$lanes timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c10_bench_grid_l$lanes.json 2> $OUT/c10_bench_grid_l$lanes.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c10_bench_bal.json 2> $OUT/c10_bench_bal.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload flat > $OUT/c10_bench_flat.json 2> $OUT/c10_bench_flat.err
tail -3 $OUT/c10_pytest.log
