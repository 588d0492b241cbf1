#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c8_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c8_pytest.log
timeout 200 python tools/probe_panel.py > $OUT/c8_panel.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c8_bench_bal.json 2> $OUT/c8_bench_bal.err
BSPB200_SOLVE_LT_WARP=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c8_bench_bal_oldlt.json 2> $OUT/c8_bench_bal_oldlt.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload stress > $OUT/c8_bench_stress.json 2> $OUT/c8_bench_stress.err
tail -3 $OUT/c8_pytest.log; cat $OUT/c8_panel.log | cut -c1-300
