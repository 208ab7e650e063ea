#!/bin/bash
# Development iteration on the GPU box: parity tests (stop at first failure), bench lines, one ncu capture of the step kernel.
# Usage: bash scripts/gpu_iter.sh [tag] [pytest -k expression]
TAG=${1:-i}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 ${2:+-k "$2"} > $OUT/pytest_gpu_$TAG.log 2>&1; tail -12 $OUT/pytest_gpu_$TAG.log
for wl in c2 c4; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_${wl}_$TAG.json').read().strip().splitlines()[-1])
    print('$wl value %.0f Mpx/s  ms/step %.4f  e2e %.0f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
except Exception as e:
    print('$wl bench failed', e); print(open('$OUT/bench_${wl}_$TAG.err').read()[-1500:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_$TAG python bench.py --workload c4 --steps 2 --warmup 3 --no-extras > $OUT/ncu_step_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 3 -c 1 -f -o $OUT/prof_step_p5_$TAG python scripts/run_once.py p5_c4 > $OUT/ncu_step_p5_$TAG.log 2>&1
