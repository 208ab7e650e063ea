"""Single-operator forward / backward launches at the C4 shape (the Executor.execute launches of the Actor): ms and fraction of the HBM peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200.functional as TF
dev = 'cuda:0'
peak, _ = bench.measured_peak_hbm()
B, H, W = 16, 2048, 3072
px = B * H * W
dgen = torch.Generator(device=dev).manual_seed(5010)
gen = torch.Generator().manual_seed(5010)
img = torch.rand(B, 3, H, W, generator=dgen, device=dev)
gout = torch.randn(B, 3, H, W, generator=dgen, device=dev)
def t(fn, n=5):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
which = [int(v) for v in sys.argv[1:]] or [6, 11, 0, 3]
for op in which:
    n = {3: 24, 5: 8, 9: 3}.get(op, 1)
    p = (torch.rand(B, n, generator=gen) * 0.5 + (0.7 if n > 1 else 0.0)).to(dev).contiguous()
    tf = t(lambda: TF._forward_raw([op], [0], img, None, 0, p, n, None, True, False, 8))
    tb = t(lambda: TF._backward_raw([op], [0], img, None, 0, p, n, gout, None, None, True, False, False, 8))
    print('op %2d  fwd %.3f ms (%.2f of peak at 24 B/px)   bwd %.3f ms (%.2f of peak at 36 B/px)' % (
        op, tf, 24 * px / tf / 1e6 / peak, tb, 36 * px / tb / 1e6 / peak))
