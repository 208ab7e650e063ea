import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner, functional as TF
NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
ex = T.Executor(T.default_options()).cuda()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
img, tgt, _ = bench.make_batch(M, 128, 128, 3010, 'cuda:0')
acc = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.time()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.time() - t0
        return r
    return w
planner.fit_params_nelder_mead = timed('fit', planner.fit_params_nelder_mead)
planner._score_outputs = timed('apply', planner._score_outputs)
orig_run = TF.DeviceNelderMead.run
def run(self, *a, **k):
    t0 = time.time(); r = orig_run(self, *a, **k); acc['nm_run'] = acc.get('nm_run', 0) + time.time() - t0
    acc.setdefault('rounds', []).append((self.P, self.S, self.rounds)); return r
TF.DeviceNelderMead.run = run
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.time()
    planner.beam_search_batch(img, tgt, ex, 8, bench.CHAIN, NAMES, 6, 1e-2)
    torch.cuda.synchronize(); print('total %.2f' % (time.time() - t0), {k: (round(v, 3) if not isinstance(v, list) else v) for k, v in acc.items()})
# GPU time per round at different stages of one step
states = img[:8].repeat(8, 1, 1, 1)[:64].contiguous()
