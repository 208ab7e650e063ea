#!/bin/bash
# Quick GPU visit: a subset of the parity tests (pytest -k expression in $1) and an optional python script ($2).
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=600 ${1:+-k "$1"} 2>&1 | tail -25
if [ -n "$2" ]; then timeout 600 python $2 2>&1 | tail -40; fi
