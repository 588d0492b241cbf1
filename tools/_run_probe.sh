timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in bal stress flat; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config4 > gpurun_out/bench_${w}_r02f.json 2> gpurun_out/bench_${w}_r02f.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_r02f.json"))
print("$w", d["ms_per_step"], d.get("factor_ms"), d.get("solve_ms"), d["value"], d["e2e"]["ms_per_step"] if d.get("e2e") else None, d["residual"], d.get("roofline_hbm",{}) and d["roofline_hbm"].get("frac"), {k:(v["launches"], round(v["ms"],3)) for k,v in d["kernel_classes"].items() if k!="timeline" and v["launches"]})
PY
done
