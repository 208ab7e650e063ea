"""Resident Nelder-Mead (t2o_nm_run_resident) against the rounds, and the per-round cost inside a cluster (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner, functional as TF
dev = 'cuda:0'
ex = T.Executor(T.default_options()).cuda()


def run(S, ops, resident, max_rounds=None, seed=5):
    img, tgt, _ = bench.make_batch(S, 128, 128, 3010 + seed, dev)
    probs = [(s, o) for s in range(S) for o in ops]
    nm = TF.DeviceNelderMead(img, tgt, [p[0] for p in probs], [p[1] for p in probs], [planner._param0(p[1], ex) for p in probs],
                             state_target=list(range(S)))
    torch.cuda.synchronize(); t0 = time.time()
    if resident:
        assert nm.run_resident(max_rounds)
        r = nm.result()
    else:
        os.environ['T2O_NM_RESIDENT'] = '0'
        r = nm.run(max_rounds=max_rounds)
    torch.cuda.synchronize(); dt = time.time() - t0
    return dt, r


for S, ops, mr in ((74, [3], 512), (74, [3, 5], 512), (148, [3, 5], 512), (512, [3, 5], 512), (64, [0, 1, 2, 3, 5, 6], None), (512, [0, 1, 2, 3, 5, 6], None)):
    for resident in (True, False):
        run(S, ops, resident, mr)
        dt, r = run(S, ops, resident, mr)
        nf = r['nfev'].numpy().reshape(S, len(ops))
        print('S=%4d ops=%s max_rounds=%s %s: %.1f ms; evals %d; per-op mean nfev %s max %s' % (
            S, ops, mr, 'resident' if resident else 'rounds  ', dt * 1e3, nf.sum(), nf.mean(0).round(0), nf.max(0)), flush=True)
        if mr is None and resident:
            q = np.sort(nf.max(1))
            print('   rounds per state: quantiles 10/50/90/99/100 %% = %s' % [int(q[int(f * (len(q) - 1))]) for f in (.1, .5, .9, .99, 1.0)])
