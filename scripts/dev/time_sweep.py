"""The full-sweep scorer launch (512 states x 168 candidates) timed alone (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import t2onet_b200.functional as TF
dev = 'cuda:0'
S, H, W = 512, 128, 128
gen = torch.Generator().manual_seed(10 + 3000)
states = torch.rand(S, 3, H, W, generator=gen).to(dev)
targets = torch.rand(64, 3, H, W, generator=gen).to(dev)
ops, prm = [], []
for op, cnt in ((0, 10), (1, 10), (2, 10), (6, 10), (5, 64), (3, 64)):
    n = {3: 24, 5: 8}.get(op, 1)
    for _ in range(cnt):
        ops.append(op)
        row = torch.zeros(24)
        row[:n] = (0.5 + torch.rand(n, generator=gen)) if n > 1 else torch.rand(1, generator=gen) * 0.5
        prm.append(row)
per = len(ops)
cb = TF.CandidateBatch(S, [s for s in range(S) for _ in range(per)], ops * S, torch.stack(prm).repeat(S, 1), dev, [s // 8 for s in range(S)])
for _ in range(3):
    TF.score_prepared(states, targets, cb)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(10):
    TF.score_prepared(states, targets, cb)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print('sweep %.3f ms  %.1f M candidates/s' % (ms, S * per / ms / 1e3))
