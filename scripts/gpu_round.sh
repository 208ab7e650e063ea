#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench lines, ncu launch list and full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $OUT/smi_$TAG.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
echo "== bench c2 (default)"; timeout 900 python bench.py > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; tail -c 2500 $OUT/bench_c2_$TAG.json; tail -3 $OUT/bench_c2_$TAG.err
echo "== bench c4"; timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-extras > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err; tail -c 1200 $OUT/bench_c4_$TAG.json; tail -3 $OUT/bench_c4_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -c 600 $OUT/bench_ref_$TAG.json
echo "== ncu launch list (c2 bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2_$TAG.csv python bench.py --steps 5 --warmup 3 --no-extras --no-graph > $OUT/ncu_c2_$TAG.log 2>&1
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ssim_kernel -s 1 -c 1 -f -o $OUT/prof_ssim_$TAG python scripts/dev/time_ssim.py > $OUT/ncu_ssim_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 2 -c 1 -f -o $OUT/prof_step_c2_$TAG python scripts/run_once.py c2 > $OUT/ncu_step_c2_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 2 -c 1 -f -o $OUT/prof_step_c4_$TAG python scripts/run_once.py bwd_c4 > $OUT/ncu_step_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fwd -s 1 -c 1 -f -o $OUT/prof_fwd_c4_$TAG python scripts/run_once.py fwd_c4 > $OUT/ncu_fwd_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -f -o $OUT/prof_score_$TAG python scripts/run_once.py score > $OUT/ncu_score_$TAG.log 2>&1
ls -la $OUT | tail -20
