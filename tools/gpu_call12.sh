#!/bin/bash
# cost-model data collection (primitive timings for the least-squares fit) + end-to-end preset sweep
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/model_fit.py collect --workloads grid,flat,flat_batch > $OUT/c12_model_fit.log 2>&1
tail -40 $OUT/c12_model_fit.log
