OUT=gpurun_out; mkdir -p $OUT
for wl in c2 c4; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 3 --no-extras > $OUT/b_$wl.json 2> $OUT/b_$wl.err; tail -2 $OUT/b_$wl.err
  python -c "
import json,sys
d=json.loads(open('$OUT/b_$wl.json').read().strip().splitlines()[-1])
print('$wl value %.0f Mpx/s  ms/step %.4f eager %.4f  e2e %.0f (%.3f ms) frac %.3f' % (d['value'], d['ms_per_step'], d['config']['eager_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c2_s5.csv python bench.py --steps 5 --warmup 3 --no-extras --no-graph > /dev/null 2>&1
grep -c step_ $OUT/launches_c2_s5.csv; grep step_ $OUT/launches_c2_s5.csv | tail -3 | cut -d, -f5,9,10,15-
