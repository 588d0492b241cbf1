#!/bin/bash
# self-validating exchange in the chained solve + unrolled gemv rows: parity tests and bench
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c15_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c15_pytest.log
tail -5 $OUT/c15_pytest.log
for wl in bal grid flat; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c15_bench_$wl.json 2> $OUT/c15_bench_$wl.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c15_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('c15_bench_')[1], 'ms', round(d['ms_per_step'],3), 'factor', round(d['factor_ms'],3), 'solve', round(d['solve_ms'],3), 'res', d['residual'], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'], d['kernel_classes']['solve_dense'])
    except Exception as e: print(f, 'ERR', e)
P
