#!/bin/bash
# fused gemv+assembleVec / split gemvT: parity tests, bench, and an ncu launch list of the GRID solve kernels
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c14_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c14_pytest.log
tail -5 $OUT/c14_pytest.log
for wl in bal grid flat; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c14_bench_$wl.json 2> $OUT/c14_bench_$wl.err
done
timeout -k 10 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload flat_batch > $OUT/c14_bench_flat_batch.json 2> $OUT/c14_bench_flat_batch.err
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv|assemble_vec|trsv|solve|invert|copy_vec" -c 3000 --csv --log-file $OUT/c14_grid_solve_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c14_ncu_grid.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c14_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('c14_bench_')[1], 'ms', round(d['ms_per_step'],3), 'factor', round(d['factor_ms'],3), 'solve', round(d['solve_ms'],3), 'res', d['residual'], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
