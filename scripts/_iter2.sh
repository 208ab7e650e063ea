OUT=gpurun_out; mkdir -p $OUT
./scripts/micro/pipe_rates > $OUT/pipe_rates.txt 2>&1; cat $OUT/pipe_rates.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
for nth in 256 192; do
  for wl in c2 c4; do
    T2O_STEP_THREADS=$nth timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nth=$nth $wl value %.0f Mpx/s  ms/step %.4f  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
  done
done
T2O_STEP_THREADS=192 timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -k "chain or fused or single" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_s3 python bench.py --workload c4 --steps 2 --warmup 3 --no-extras > $OUT/ncu_step_c4_s3.log 2>&1
T2O_STEP_THREADS=192 timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 3 -c 1 -f -o $OUT/prof_step_c4_s3_192 python bench.py --workload c4 --steps 2 --warmup 3 --no-extras > $OUT/ncu_step_c4_s3_192.log 2>&1
