"""One Nelder-Mead round's scorer launch in isolation: S states x k candidates (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import t2onet_b200.functional as TF
dev = 'cuda:0'
def t(fn, n=50):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
for S in (8, 64, 512):
    states = torch.rand(S, 3, 128, 128, device=dev); targets = torch.rand(max(S // 8, 1), 3, 128, 128, device=dev)
    st_t = [s // 8 for s in range(S)]
    for ops in ([3], [0, 1, 2, 3, 5, 6]):
        cand_state = [s for s in range(S) for _ in ops]
        cb = TF.CandidateBatch(S, cand_state, ops * S, torch.rand(S * len(ops), 24) + 0.5, dev, st_t)
        g = torch.cuda.CUDAGraph()
        TF.score_prepared(states, targets, cb); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(20): TF.score_prepared(states, targets, cb)
        print('S=%4d  %d candidates/state: %.1f us per launch (graph of 20)' % (S, len(ops), t(lambda: g.replay(), 10) / 20))
