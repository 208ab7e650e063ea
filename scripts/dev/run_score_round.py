"""One configuration of the Nelder-Mead-round scorer for profiler captures: python scripts/dev/run_score_round.py [S] [ncand]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import t2onet_b200.functional as TF
dev = 'cuda:0'
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ops = [3, 5, 0, 1, 2, 6][:k]
states = torch.rand(S, 3, 128, 128, device=dev); targets = torch.rand(max(S // 8, 1), 3, 128, 128, device=dev)
st_t = [s // 8 for s in range(S)]
cand_state = [s for s in range(S) for _ in ops]
cb = TF.CandidateBatch(S, cand_state, ops * S, torch.rand(S * len(ops), 24) + 0.5, dev, st_t)
for _ in range(6):
    TF.score_prepared(states, targets, cb)
torch.cuda.synchronize()
print('done')
