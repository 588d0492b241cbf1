set -x
python bench.py > gpurun_out/bench_bal_r02g.json 2> gpurun_out/bench_bal_r02g.err
python bench.py --impl reference > gpurun_out/bench_reference_r02g.json 2> gpurun_out/bench_reference_r02g.err
for w in grid stress flat bal_small; do python bench.py --workload $w --steps 5 --warmup 3 --no-config4 > gpurun_out/bench_${w}_r02g.json 2> gpurun_out/bench_${w}_r02g.err; done
python tools/check_lumpchol.py > gpurun_out/check_lumpchol_r02g.txt 2>&1
python tools/probe_lumpchol.py > gpurun_out/probe_lumpchol_r02g.txt 2>&1
python tools/probe_elim.py > gpurun_out/probe_elim_r02g.txt 2>&1
BSPB200_PROFILE_TIMELINE=1 python bench.py --workload grid --steps 3 --warmup 2 --no-cpu-baseline --no-ref-cuda --no-config4 > gpurun_out/grid_timeline_r02g.json 2>/dev/null
python tools/timeline_summary.py gpurun_out/grid_timeline_r02g.json 0.25 > gpurun_out/grid_timeline_r02g.txt
tail -c 1500 gpurun_out/bench_bal_r02g.json
