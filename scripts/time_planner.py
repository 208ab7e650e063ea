"""Wall-clock of the drop-in beam_search on one GPU (development aid): python scripts/time_planner.py [pairs] [beam]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import t2onet_b200 as T
from t2onet_b200 import planner

NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
beam = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ex = T.Executor(T.default_options()).cuda()
img, tgt, _ = bench.make_batch(npairs, 128, 128, 3010, 'cuda:0')
for i in range(npairs):
    cnt = [0]
    torch.cuda.synchronize(); t0 = time.time()
    actions, Is = planner.beam_search(img[i:i + 1], tgt[i:i + 1], None, ex, None, beam, bench.CHAIN, NAMES, 6, 1e-2, 'L1',
                                      'Nelder-Mead', counter=cnt)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('pair %d: %.2f s, %d candidates scored (%.0f /s), best %s dist %.4f' % (
        i, dt, cnt[0], cnt[0] / dt, [a[0] for a in actions[0]], actions[0][-1][2] if actions[0] else -1))

M = int(sys.argv[3]) if len(sys.argv) > 3 else 32
img, tgt, _ = bench.make_batch(M, 128, 128, 3010, 'cuda:0')
for rep in range(2):
    cnt = [0]
    torch.cuda.synchronize(); t0 = time.time()
    res = planner.beam_search_batch(img, tgt, ex, beam, bench.CHAIN, NAMES, 6, 1e-2, counter=cnt)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('batch of %d pairs: %.2f s (%.3f s/pair), %d candidates scored (%.0f /s)' % (M, dt, dt / M, cnt[0], cnt[0] / dt))
