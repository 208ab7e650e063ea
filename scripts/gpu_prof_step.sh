#!/bin/bash
# ncu full capture of the fused step kernel at the C4 and C2 shapes.  Usage: bash scripts/gpu_prof_step.sh [tag]
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_$TAG python bench.py --workload c4 --steps 2 --warmup 3 --no-extras > $OUT/ncu_step_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c2_$TAG python bench.py --workload c2 --steps 2 --warmup 3 --no-extras > $OUT/ncu_step_c2_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 3 -c 1 -f -o $OUT/prof_step_p5_$TAG python scripts/run_once.py p5_c4 > $OUT/ncu_step_p5_$TAG.log 2>&1
ls -la $OUT | tail -5
