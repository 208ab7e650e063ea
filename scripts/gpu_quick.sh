#!/bin/bash
# Quick GPU visit: parity tests, then the two bench lines.  Usage: bash scripts/gpu_quick.sh [tag]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > $OUT/pytest_gpu_$TAG.log 2>&1; tail -15 $OUT/pytest_gpu_$TAG.log
echo "== bench c2"; timeout 600 python bench.py --no-extras > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench_c2_$TAG.json').read().strip().splitlines()[-1])
print('c2 value %.0f Mpx/s  ms/step %.4f  e2e %.0f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"; tail -3 $OUT/bench_c2_$TAG.err
echo "== bench c4"; timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-extras > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench_c4_$TAG.json').read().strip().splitlines()[-1])
print('c4 value %.0f Mpx/s  ms/step %.4f  e2e %.0f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"; tail -3 $OUT/bench_c4_$TAG.err
