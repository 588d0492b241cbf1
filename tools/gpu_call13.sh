#!/bin/bash
# flag-chained dense triangular solve: parity tests, then the bench workloads with the chain on / off
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain_solve" ) > $OUT/c13_pytest_chain.log 2>&1; echo "pytest exit $?" >> $OUT/c13_pytest_chain.log
tail -5 $OUT/c13_pytest_chain.log
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c13_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c13_pytest.log
tail -5 $OUT/c13_pytest.log
for wl in bal grid flat; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c13_bench_$wl.json 2> $OUT/c13_bench_$wl.err
BSPB200_CHAIN_SOLVE=0 timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c13_bench_${wl}_nochain.json 2> $OUT/c13_bench_${wl}_nochain.err
done
timeout -k 10 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload flat_batch > $OUT/c13_bench_flat_batch.json 2> $OUT/c13_bench_flat_batch.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c13_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('c13_bench_')[1], 'ms', round(d['ms_per_step'],3), 'factor', round(d['factor_ms'],3), 'solve', round(d['solve_ms'],3), 'res', d['residual'], 'e2e', round(d['e2e']['ms_per_step'],2))
    except Exception as e: print(f, 'ERR', e)
P
