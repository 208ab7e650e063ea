#!/bin/bash
# Kernel iteration on the GPU box: operator / step parity tests, bench lines (c4, c2), ncu captures of the fused step.  bash scripts/gpu_kern.sh [tag]
TAG=${1:-k}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rows.py tests/test_gpu_nm.py tests/test_gpu_metrics.py tests/test_gpu_convert.py -m gpu -q -x --timeout=600 > $OUT/pytest_kern_$TAG.log 2>&1; tail -6 $OUT/pytest_kern_$TAG.log
for wl in c4 c2; do
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_$TAG python bench.py --workload c4 --steps 2 --warmup 3 --no-extras --no-graph > $OUT/ncu_step_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 12 -c 1 -f -o $OUT/prof_step_c2_$TAG python bench.py --workload c2 --steps 2 --warmup 3 --no-extras --no-graph > $OUT/ncu_step_c2_$TAG.log 2>&1
ls -la $OUT/*$TAG* | tail -8
