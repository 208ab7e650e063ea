"""Per-phase clocks of the resident Nelder-Mead kernel (library built with T2O_NVCC_EXTRA=-DT2O_RES_PROBE; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import t2onet_b200 as T
from t2onet_b200 import planner, functional as TF, _lib
dev = 'cuda:0'
ex = T.Executor(T.default_options()).cuda()
for S, ops, mr in ((64, [3], 600), (64, [3, 5], 600), (64, [0, 1, 2, 3, 5, 6], None), (512, [3, 5], 600)):
    img, tgt, _ = bench.make_batch(S, 128, 128, 3015, dev)
    probs = [(s, o) for s in range(S) for o in ops]
    nm = TF.DeviceNelderMead(img, tgt, [p[0] for p in probs], [p[1] for p in probs], [planner._param0(p[1], ex) for p in probs],
                             state_target=list(range(S)))
    assert nm.run_resident(mr)
    torch.cuda.synchronize()
    ws = _lib.workspace(torch.device(dev), 1 << 19)
    pr = ws[(1 << 18):(1 << 18) + 2 * 8 * 6 * 8].view(torch.int64).cpu().view(2, 8, 6)
    nmp = ws[(1 << 18) + 128 * 8:(1 << 18) + 140 * 8].view(torch.int64).cpu().tolist()
    print('S=%d ops=%s   NM stages (cumulative cycles, count last):' % (S, ops), nmp)
    for rank in range(2):
        for w in range(8):
            c = pr[rank, w].tolist()
            n = max(c[5], 1)
            print('  rank %d warp %d: rounds %5d  cycles/round: syncA %6.0f tables %6.0f arith %6.0f syncB %6.0f advance %6.0f  total %6.0f' % (
                rank, w, c[5], c[0] / n, c[1] / n, c[2] / n, c[3] / n, c[4] / n, sum(c[:5]) / n))
