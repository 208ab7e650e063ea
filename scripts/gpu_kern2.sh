#!/bin/bash
# Kernel iteration with the experimental step variant: bash scripts/gpu_kern2.sh [tag]
TAG=${1:-k}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rows.py -m gpu -q -x --timeout=600 > $OUT/pytest_kern_$TAG.log 2>&1; tail -3 $OUT/pytest_kern_$TAG.log
echo "== parity tests, variant 2"; T2O_STEP_VARIANT=2 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=600 > $OUT/pytest_kern_v2_$TAG.log 2>&1; tail -3 $OUT/pytest_kern_v2_$TAG.log
for var in 0 2; do
for wl in c4 c2; do
  T2O_STEP_VARIANT=$var timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras > $OUT/bench_${wl}_v${var}_$TAG.json 2> $OUT/bench_${wl}_v${var}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_${wl}_v${var}_$TAG.json').read().strip().splitlines()[-1])
    print('variant $var $wl value %.0f Mpx/s  ms/step %.4f  e2e %.0f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
except Exception as e:
    print('$wl bench failed', e); print(open('$OUT/bench_${wl}_v${var}_$TAG.err').read()[-1500:])
PY
done
T2O_STEP_VARIANT=$var timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_v${var}_$TAG python bench.py --workload c4 --steps 2 --warmup 3 --no-extras --no-graph > $OUT/ncu_step_c4_v${var}_$TAG.log 2>&1
done
ls -la $OUT/*$TAG* | tail -8
