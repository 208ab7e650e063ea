#!/bin/bash
# One GPU-box visit of round 2: parity tests, smoke, default bench line, reference arm.  bash scripts/gpu_visit.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $OUT/smi_$TAG.txt
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --durations=12 -s > $OUT/pytest_gpu_$TAG.log 2>&1; tail -40 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 1500 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -c 900 $OUT/bench_ref_$TAG.json; tail -3 $OUT/bench_ref_$TAG.err
