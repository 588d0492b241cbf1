timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in bal grid flat; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w', round(d['ms_per_step'],3), round(d['factor_ms'],3), round(d['solve_ms'],3), d['residual'], {k:(v['launches'], round(v['ms'],3)) for k,v in d['kernel_classes'].items() if k!='timeline' and v['launches']})"; done
