OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; tail -3 $OUT/bench_n2.err; cat $OUT/bench_n2.json | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; cat $OUT/bench_ref_n2.json | cut -c1-400
( time python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real; tail -2 $OUT/bench_default.err; cat $OUT/bench_default.json | cut -c1-3000
