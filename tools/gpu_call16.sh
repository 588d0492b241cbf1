#!/bin/bash
# chain with back-off + prefetched rhs + rows-below fused into the forward launch; 16-byte operand loads in the gather
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c16_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c16_pytest.log
tail -5 $OUT/c16_pytest.log
for wl in bal grid flat stress; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c16_bench_$wl.json 2> $OUT/c16_bench_$wl.err
done
BSPB200_GATHER_VEC16=0 timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bal > $OUT/c16_bench_bal_novec.json 2> $OUT/c16_bench_bal_novec.err
BSPB200_CHAIN_FUSE_GEMV=0 timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c16_bench_grid_nofuse.json 2> $OUT/c16_bench_grid_nofuse.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c16_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f.split('c16_bench_')[1], 'ms', round(d['ms_per_step'],3), 'factor', round(d['factor_ms'],3), 'solve', round(d['solve_ms'],3), 'res', d['residual'], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'], 'solve_dense', round(k['solve_dense']['ms'],3), 'gather', round(k['elim_gather']['ms'],3))
    except Exception as e: print(f, 'ERR', e)
P
