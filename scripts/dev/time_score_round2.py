"""Scorer launch time over (states, live candidates per state) for both Nelder-Mead-round kernels (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import t2onet_b200.functional as TF
dev = 'cuda:0'
def t(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for S in (32, 64, 128, 192, 256, 384, 512):
    states = torch.rand(S, 3, H, H, device=dev); targets = torch.rand(max(S // 8, 1), 3, H, H, device=dev)
    st_t = [s // 8 for s in range(S)]
    row = []
    for live in (1, 2, 6):
        ops = [3, 5, 0, 1, 2, 6]
        cand_op = []
        for s in range(S):
            cand_op += [o if i < live else -2 for i, o in enumerate(ops)]       # 6 fits per state, `live` of them unfinished
        cand_state = [s for s in range(S) for _ in ops]
        cb = TF.CandidateBatch(S, cand_state, cand_op, torch.rand(S * 6, 24) + 0.5, dev, st_t)
        res = []
        for env in ('0', '1'):
            os.environ['T2O_SCORE_STREAM'] = env
            TF.score_prepared(states, targets, cb); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20): TF.score_prepared(states, targets, cb)
            res.append(t(lambda: g.replay(), 5) / 20)
        row.append('live %d: old %.1f stream %.1f' % (live, res[0], res[1]))
    print('S=%4d  ' % S + '   '.join(row))
