#!/bin/bash
# flag protocol restored (+ prefetched rhs); rows-below fusion as opt-in variant
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c17_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c17_pytest.log
tail -5 $OUT/c17_pytest.log
for wl in bal grid flat; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c17_bench_$wl.json 2> $OUT/c17_bench_$wl.err
done
BSPB200_CHAIN_FUSE_GEMV=1 timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c17_bench_grid_fuse.json 2> $OUT/c17_bench_grid_fuse.err
BSPB200_CHAIN_FUSE_GEMV=1 timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grid or solve_many or partial" > $OUT/c17_pytest_fuse.log 2>&1; tail -2 $OUT/c17_pytest_fuse.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c17_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f.split('c17_bench_')[1], 'ms', round(d['ms_per_step'],3), 'factor', round(d['factor_ms'],3), 'solve', round(d['solve_ms'],3), 'res', d['residual'], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'], 'solve_dense', round(k['solve_dense']['ms'],3), 'gather', round(k['elim_gather']['ms'],3))
    except Exception as e: print(f, 'ERR', e)
P
