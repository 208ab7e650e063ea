"""cProfile of one planner.beam_search_batch call (development aid)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner
NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
ex = T.Executor(T.default_options()).cuda()
img, tgt, _ = bench.make_batch(64, 128, 128, 3010, 'cuda:0')
planner.beam_search_batch(img, tgt, ex, 8, bench.CHAIN, NAMES, 6, 1e-2)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
planner.beam_search_batch(img, tgt, ex, 8, bench.CHAIN, NAMES, 6, 1e-2)
torch.cuda.synchronize()
pr.disable()
print('total %.3f s' % (time.time() - t0))
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
