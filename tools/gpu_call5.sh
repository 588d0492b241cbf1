#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/c5_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/c5_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c5_bench_bal.json 2> $OUT/c5_bench_bal.err
mk() { python -c "
f=$1
potrf=[1.0e-05*f,4e-07,0.0,2.0e-14]
trsm=[3.0e-06*f,0,0,1.0e-09,0,5.0e-14]
syge=[1.2e-05*f,0,0,2.0e-09,0,6.0e-14]
asm=[1.0e-05*f,1.0e-09,1.0e-09,1.0e-11]
print(','.join(str(x) for x in potrf+trsm+syge+asm))"; }
for f in 0.1 0.03; do
  BSPB200_MODEL_PARAMS=$(mk $f) timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid --model 3 > $OUT/c5_bench_grid_f$f.json 2> $OUT/c5_bench_grid_f$f.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload grid --model 1 > $OUT/c5_bench_grid_m1.json 2> $OUT/c5_bench_grid_m1.err
BSPB200_PDL=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_f64_kernelILi128 -s 27 -c 9 -o $OUT/c5_prof_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/c5_ncu_gemm.log 2>&1
BSPB200_PDL=0 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:panel2_kernel|elim_gather|elim_factor|solve_step_inv' -c 10 -o $OUT/c5_prof_misc \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/c5_ncu_misc.log 2>&1
ls -la $OUT/*.ncu-rep; tail -3 $OUT/c5_pytest.log
