#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
BSPB200_GEMM_CFG=1 timeout 300 python tools/probe2.py > $OUT/c9_probe2_cfg1.log 2>&1; cp $OUT/probe2.json $OUT/c9_probe2_cfg1.json
BSPB200_GEMM_CFG=3 timeout 300 python tools/probe2.py > $OUT/c9_probe2_cfg3.log 2>&1; cp $OUT/probe2.json $OUT/c9_probe2_cfg3.json
BSPB200_GEMM_CFG=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c9_bench_bal_cfg3.json 2> $OUT/c9_bench_bal_cfg3.err
BSPB200_GEMM_CFG=3 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c9_bench_grid_cfg3.json 2> $OUT/c9_bench_grid_cfg3.err
tail -2 $OUT/c9_probe2_cfg1.log; tail -2 $OUT/c9_probe2_cfg3.log
