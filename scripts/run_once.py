"""Tiny drivers for profiler captures: python scripts/run_once.py {fwd_c4|bwd_c4|score|c2|resident}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import t2onet_b200.functional as TF  # noqa: E402

what = sys.argv[1]
dev = 'cuda:0'
if what in ('fwd_c4', 'bwd_c4'):
    img, tgt, params = bench.make_batch(16, 2048, 3072, 4010, dev)
    for _ in range(3):
        if what == 'fwd_c4':
            TF._forward_raw(bench.CHAIN, [0, 1, 2, 3, 27, 35], img, None, 0, torch.cat(params, 1).contiguous(), 36, tgt, True, True, 8)
        else:
            TF.chain_forward_backward(img, bench.CHAIN, params, tgt)
elif what == 'p5_c4':
    img, tgt, params = bench.make_batch(16, 2048, 3072, 4010, dev)
    for _ in range(4):
        TF.chain_forward_backward(img, bench.CHAIN[:5], params[:5], tgt)
elif what == 'c2':
    img, tgt, params = bench.make_batch(64, 128, 128, 2010, dev)
    for _ in range(3):
        TF.chain_forward_backward(img, bench.CHAIN, params, tgt)
elif what == 'score':
    S, H, W = 512, 128, 128
    gen = torch.Generator().manual_seed(3010)
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
    cb = TF.CandidateBatch(S, [s for s in range(S) for _ in ops], ops * S, torch.stack(prm).repeat(S, 1), dev,
                           [s // 8 for s in range(S)])
    for _ in range(3):
        TF.score_prepared(states, targets, cb)
torch.cuda.synchronize()
print('done', what)

if what == 'resident':
    # every fit of a planner step in one launch: 132 states x the six FiveK operators (two waves of clusters)
    import t2onet_b200 as T
    from t2onet_b200 import planner
    ex = T.Executor(T.default_options()).cuda()
    S = 132
    img, tgt, _ = bench.make_batch(S, 128, 128, 3015, dev)
    probs = [(s, o) for s in range(S) for o in bench.CHAIN]
    nm = TF.DeviceNelderMead(img, tgt, [p[0] for p in probs], [p[1] for p in probs], [planner._param0(p[1], ex) for p in probs],
                             state_target=list(range(S)))
    assert nm.run_resident()
    torch.cuda.synchronize()
    r = nm.result()
    print('evaluations', int(r['nfev'].sum()), 'done', bool(r['done'].all()))
torch.cuda.synchronize()
