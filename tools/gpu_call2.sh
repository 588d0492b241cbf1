#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/probe2.py > $OUT/c2_probe2.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c2_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c2_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c2_bench_bal_staged.json 2> $OUT/c2_bench_bal_staged.err
BSPB200_GATHER=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c2_bench_bal_direct.json 2> $OUT/c2_bench_bal_direct.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload stress > $OUT/c2_bench_stress_staged.json 2> $OUT/c2_bench_stress_staged.err
tail -3 $OUT/c2_pytest.log; tail -5 $OUT/c2_probe2.log
