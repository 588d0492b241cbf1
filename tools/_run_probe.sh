BSPB200_PROFILE_TIMELINE=1 timeout 300 python bench.py --workload grid --steps 3 --warmup 2 --no-cpu-baseline --no-ref-cuda --no-config4 > gpurun_out/grid_timeline.json 2>/dev/null
python tools/timeline_summary.py gpurun_out/grid_timeline.json 0.25
