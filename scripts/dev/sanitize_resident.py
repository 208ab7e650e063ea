"""A short resident Nelder-Mead launch for compute-sanitizer (memcheck / racecheck): a few states, two tiles per image, masks,
and a single-tile shape; 60 rounds each (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner, functional as TF
dev = 'cuda:0'
ex = T.Executor(T.default_options()).cuda()
SHAPES = ((5, 64, 96, True), (3, 32, 32, False), (3, 128, 128, False)) if len(sys.argv) < 2 else ((5, 64, 96, True), (2, 128, 128, False))
ROUNDS = 60 if len(sys.argv) < 2 else 24
for S, H, W, masked in SHAPES:
    img, tgt, _ = bench.make_batch(S, H, W, 11, dev)
    probs = [(s, o) for s in range(S) for o in bench.CHAIN]
    kw = {}
    if masked:
        g = torch.Generator().manual_seed(3)
        kw = dict(masks=(torch.rand(2, 1, H, W, generator=g) > 0.4).float().cuda(), prob_mask=[(i % 3) - 1 for i in range(len(probs))])
    nm = TF.DeviceNelderMead(img, tgt, [p[0] for p in probs], [p[1] for p in probs], [planner._param0(p[1], ex) for p in probs],
                             state_target=list(range(S)), **kw)
    assert nm.run_resident(ROUNDS)
    torch.cuda.synchronize()
    print(S, H, W, masked, 'evaluations', int(nm.result()['nfev'].sum()))
