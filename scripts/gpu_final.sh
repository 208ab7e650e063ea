#!/bin/bash
# Round-end GPU visit: parity tests, smoke, the bench lines (default = C4 + C2 + planner, reference arm, C3, C1), the ncu launch
# list of the default bench command.  Usage (under gpurun, from the repo root):  bash scripts/gpu_final.sh [tag]
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $OUT/smi_$TAG.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
echo "== bench default"; ( time timeout 900 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err ) 2>&1 | grep real; tail -2 $OUT/bench_default_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -c 500 $OUT/bench_ref_$TAG.json
echo "== bench c3"; timeout 600 python bench.py --workload c3 > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; tail -c 700 $OUT/bench_c3_$TAG.json
echo "== bench c2 / c1"; for wl in c2 c1; do timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err; done
echo "== ncu launch list (default bench command, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_default_$TAG.csv python bench.py --steps 5 --warmup 3 --no-extras --no-graph > $OUT/ncu_default_$TAG.log 2>&1
tail -3 $OUT/launches_default_$TAG.csv | cut -c1-300
python - <<PY
import json
d=json.loads(open("$OUT/bench_default_$TAG.json").read().strip().splitlines()[-1])
print("value %.0f Mpx/s  ms/step %.4f  e2e %.0f  frac %.3f  issue %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["issue"]["frac"]))
print("c2", d["c2"]["value"], d["c2"]["ms_per_step"], d["c2"]["roofline"]["frac"])
p=d["planner"]; print("planner", p["value"], p["e2e"]["pairs_per_s"], p["e2e"]["seconds"], "gier", p["gier"].get("pairs_per_s"), "cpu", p["cpu_baseline"].get("value"))
print("cpu_baseline", d["cpu_baseline"])
PY
