#!/bin/bash
# 2-GPU validation of the multi-rank bench path (weak scaling on the headline workload, strong scaling on the batch)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/c11_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c11_bench_bal_n2.json 2> $OUT/c11_bench_bal_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --workload flat_batch > $OUT/c11_bench_flat_batch_n2.json 2> $OUT/c11_bench_flat_batch_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/c11_bench_ref_n2.json 2> $OUT/c11_bench_ref_n2.err
BSPB200_LANES=16 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid > $OUT/c11_bench_grid_l16.json 2> $OUT/c11_bench_grid_l16.err
head -c 400 $OUT/c11_bench_bal_n2.json; echo; head -c 300 $OUT/c11_bench_flat_batch_n2.json; echo; tail -3 $OUT/c11_bench_bal_n2.err
