OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -4
python - <<'PY'
import torch, bench, t2onet_b200.functional as TF
dev='cuda:0'
img, tgt, params = bench.make_batch(16, 2048, 3072, 4010, dev)
packed = torch.cat(params, 1).contiguous()
def t(fn, n=10):
    for _ in range(3): fn()
    s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
px = 16*2048*3072
ms = t(lambda: TF._forward_raw(bench.CHAIN, [0,1,2,3,27,35], img, None, 0, packed, 36, tgt, True, True, 8))
print('fwd6 out+l1   %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36*px/ms/1e6))
ms = t(lambda: TF._forward_raw(bench.CHAIN, [0,1,2,3,27,35], img, None, 0, packed, 36, tgt, False, True, 8))
print('fwd6 l1 only  %.3f ms  %.0f GB/s (24 B/px)' % (ms, 24*px/ms/1e6))
ms = t(lambda: TF._forward_raw(bench.CHAIN[:5], [0,1,2,3,27], img, None, 0, packed, 36, tgt, True, True, 8))
print('fwd5 flat out+l1 %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36*px/ms/1e6))
ms = t(lambda: TF._forward_raw([6], [35], img, None, 0, packed, 36, tgt, True, True, 8))
print('sharp only out+l1 %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36*px/ms/1e6))
ms = t(lambda: TF.l1_sum(img, tgt))
print('l1_sum        %.3f ms  %.0f GB/s (24 B/px)' % (ms, 24*px/ms/1e6))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fwd_rows -s 1 -c 1 -f -o $OUT/prof_fwd_rows_c4_s7 python scripts/run_once.py fwd_c4 > $OUT/ncu_fwd_rows_s7.log 2>&1
