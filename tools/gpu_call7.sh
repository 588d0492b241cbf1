#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for uf in 12 6 4 3 2; do BSPB200_PANEL_UF=$uf timeout 200 python tools/probe_panel.py >> $OUT/c7_panel_uf.log 2>&1; done
( BSPB200_GATHER=3 timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c7_pytest_coop.log 2>&1; echo "pytest exit $?" >> $OUT/c7_pytest_coop.log
BSPB200_GATHER=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c7_bench_bal_coop.json 2> $OUT/c7_bench_bal_coop.err
BSPB200_GATHER=3 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload stress > $OUT/c7_bench_stress_coop.json 2> $OUT/c7_bench_stress_coop.err
cat $OUT/c7_panel_uf.log | cut -c1-400; tail -3 $OUT/c7_pytest_coop.log
