import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner
NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
ex = T.Executor(T.default_options()).cuda()
M = 8
img, tgt, _ = bench.make_batch(M, 128, 128, 3010, 'cuda:0')
runs = []
for rep in range(3):
    cnt = [0]
    res = planner.beam_search_batch(img, tgt, ex, 8, bench.CHAIN, NAMES, 6, 1e-2, counter=cnt)
    runs.append(res)
    print('run', rep, cnt[0], [len(r[0][0]) for r in res], [round(r[0][0][-1][2], 6) for r in res])
for m in range(M):
    a0 = [[(a[0], a[2]) for a in seq] for seq in runs[0][m][0]]
    for rep in (1, 2):
        a1 = [[(a[0], a[2]) for a in seq] for seq in runs[rep][m][0]]
        if a0 != a1:
            print('pair', m, 'run', rep, 'differs'); print(a0[0]); print(a1[0]); break
# fits determinism on one step
states = img[:4].contiguous()
problems = [(s, op) for s in range(4) for op in bench.CHAIN]
outs = []
for rep in range(3):
    fits = planner.fit_params_nelder_mead(states, tgt[:4].contiguous(), problems, ex, state_target=[0, 1, 2, 3])
    outs.append([(f.nfev, f.nit, f.fun) for f in fits])
    print('fits', rep, sum(f.nfev for f in fits))
print('fits equal', outs[0] == outs[1] == outs[2])
for a, b, pr in zip(outs[0], outs[1], problems):
    if a != b: print(pr, a, b)
