#!/bin/bash
# end-of-session evidence: gpu tests, smoke, bench lines of every config, both baselines (CPU port, restated reference
# CUDA backend), ncu launch list of the headline step and one full capture of the chained-solve kernel
OUT=gpurun_out; mkdir -p $OUT
( timeout -k 10 600 python -m pytest tests -m gpu -x -q ) > $OUT/c18_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c18_pytest.log
tail -4 $OUT/c18_pytest.log
( timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/c18_smoke.log 2>&1; tail -1 $OUT/c18_smoke.log
timeout -k 10 400 python bench.py --steps 10 --warmup 3 > $OUT/c18_bench_bal.json 2> $OUT/c18_bench_bal.err
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/c18_bench_reference.json 2> $OUT/c18_bench_reference.err
for wl in bal grid flat; do
timeout -k 10 200 python bench.py --impl ref_cuda --steps 3 --warmup 2 --workload $wl > $OUT/c18_bench_refcuda_$wl.json 2> $OUT/c18_bench_refcuda_$wl.err
done
for wl in grid flat stress bal_small; do
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $OUT/c18_bench_$wl.json 2> $OUT/c18_bench_$wl.err
done
timeout -k 10 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload flat_batch > $OUT/c18_bench_flat_batch.json 2> $OUT/c18_bench_flat_batch.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $OUT/c18_launches_bal.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/c18_ncu_launches.log 2>&1
timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:trsv_chain -s 2 -c 2 -o $OUT/c18_trsv_chain python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/c18_ncu_chain.log 2>&1
ncu -i $OUT/c18_trsv_chain.ncu-rep --page raw --csv > $OUT/c18_trsv_chain_raw.csv 2>/dev/null
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c18_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('c18_bench_')[1], d.get('impl','b200'), 'ms', round(d['ms_per_step'],3), 'GF/s', round(d['value'],1), 'factor', round(d.get('factor_ms',0),3), 'solve', round(d.get('solve_ms',0),3), 'x', d.get('x_head'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
