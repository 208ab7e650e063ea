import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import t2onet_b200 as T
from oracle import ops as O
rec = json.load(open(os.path.join(ROOT, 'tests/golden/planner_pairs.json')))
d = np.load(os.path.join(ROOT, 'tests/golden/planner_pairs.npz'))
I0, Igt = torch.from_numpy(d['I0']).cuda(), torch.from_numpy(d['Igt']).cuda()
st = rec['settings']
ex = T.Executor(T.default_options()).cuda()
res = T.planner.beam_search_batch(I0, Igt, ex, st['beam'], st['operations'], O.ACTION_NAMES, st['max_step'], st['err'])
for m, (pair, (actions, Is)) in enumerate(zip(rec['pairs'], res)):
    ref = [[(a[0][:4], round(a[2], 5)) for a in seq] for seq in pair['actions']]
    our = [[(a[0][:4], round(a[2], 5)) for a in seq] for seq in actions]
    same = [a[0] for a in ref[0]] == [a[0] for a in our[0]]
    if not same:
        print('pair', m, 'planted', pair['planted'], 'init %.5f' % pair['init_dist'])
        for b in range(len(ref)):
            print('   ref', ref[b]); print('   our', our[b] if b < len(our) else None)
