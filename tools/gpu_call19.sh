#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout -k 5 70 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/c19_bench_bal.json 2> $OUT/c19_bench_bal.err
python -c "
import json;d=json.loads(open('gpurun_out/c19_bench_bal.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['clocks'])"
tail -3 $OUT/c19_bench_bal.err
