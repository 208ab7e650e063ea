#!/bin/bash
# GPU-box visit: the parity tests only (verbose on the planner / actor evidence).  bash scripts/gpu_tests.sh [tag] [pytest args]
TAG=${1:-r02}
shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $OUT/smi_$TAG.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --durations=15 "$@" > $OUT/pytest_gpu_$TAG.log 2>&1
tail -60 $OUT/pytest_gpu_$TAG.log
