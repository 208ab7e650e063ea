import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from t2onet_b200 import metrics as MT
def t(fn, n=5):
    for _ in range(2): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
a = torch.rand(16, 3, 2048, 3072, device='cuda'); b = torch.rand_like(a)
ms = t(lambda: MT.ssim_sum(a, b)); px = 16 * 2048 * 3072
print('SSIM C4 shape %.3f ms  %.0f GB/s (24 B/px)' % (ms, 24 * px / ms / 1e6))
