timeout 300 python tools/check_lumpchol.py 426,300 600,5000 2000,2 5226,0 | tail -4
PROBE_UFS=0 timeout 300 python tools/probe_lumpchol.py 2>&1 | grep -E "ms \(min|stage_L1|syrk|chain step"
