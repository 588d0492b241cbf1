#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c6_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c6_pytest.log
timeout 600 python tools/probe2.py > $OUT/c6_probe2.log 2>&1; cp $OUT/probe2.json $OUT/c6_probe2.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c6_bench_bal.json 2> $OUT/c6_bench_bal.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c6_bench_grid.json 2> $OUT/c6_bench_grid.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload flat > $OUT/c6_bench_flat.json 2> $OUT/c6_bench_flat.err
timeout 500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload flat_batch > $OUT/c6_bench_flat_batch.json 2> $OUT/c6_bench_flat_batch.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_f64_kernel -s 178 -c 13 -o $OUT/c6_prof_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/c6_ncu_gemm.log 2>&1
ls -la $OUT/*.ncu-rep; tail -3 $OUT/c6_pytest.log; grep -E "panel_n96_rows5130" -A3 $OUT/c6_probe2.log | head
