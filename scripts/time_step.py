"""Quick timings of the step kernels on one GPU (development aid): python scripts/time_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import t2onet_b200.functional as TF

dev = 'cuda:0'
def t(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n

img, tgt, params = bench.make_batch(16, 2048, 3072, 4010, dev)
px = 16 * 2048 * 3072
gout = torch.randn_like(img)
print('C4 fused 6-op step      %.3f ms' % t(lambda: TF.chain_forward_backward(img, bench.CHAIN, params, tgt)))
print('C4 fused 5-op flat step %.3f ms' % t(lambda: TF.chain_forward_backward(img, bench.CHAIN[:5], params[:5], tgt)))
print('C4 sharp-first 6-op     %.3f ms' % t(lambda: TF.chain_forward_backward(img, [6, 0, 1, 2, 3, 5], [params[5]] + params[:5], tgt)))
packed6 = torch.cat(params, 1).contiguous()
offs6 = [0, 1, 2, 3, 27, 35]
ms = t(lambda: TF._forward_raw(bench.CHAIN, offs6, img, None, 0, packed6, packed6.shape[1], tgt, True, True, 8))
print('C4 forward 6-op + L1    %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36 * px / ms / 1e6))
p6 = params[5].contiguous()
ms = t(lambda: TF._forward_raw([6], [0], img, None, 0, p6, 1, tgt, True, True, 8))
print('C4 sharpness fwd + L1   %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36 * px / ms / 1e6))
ms = t(lambda: TF._backward_raw([6], [0], img, None, 0, p6, 1, gout, None, None, True, False, False, 8))
print('C4 sharpness bwd        %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36 * px / ms / 1e6))
ms = t(lambda: TF._backward_raw([6], [0], img, None, 0, p6, 1, None, tgt, bench.fused_scale(16, 2048, 3072, dev), False, True, True, 8))
print('C4 sharpness fused step %.3f ms  %.0f GB/s (36 B/px)' % (ms, 36 * px / ms / 1e6))
del img, tgt, gout
img, tgt, params = bench.make_batch(64, 128, 128, 2010, dev)
packed = torch.cat(params, 1).contiguous()
fs = TF.FusedStep(bench.CHAIN, 64, 128, 128, dev, reuse_outputs=True)
g = torch.cuda.CUDAGraph()
fs(img, packed, tgt); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(20): fs(img, packed, tgt)
print('C2 fused 6-op step      %.2f us (graph of 20, L2-resident)' % (t(lambda: g.replay(), 5) / 20 * 1e3))
