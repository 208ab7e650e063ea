"""Record planner transcripts of the UNMODIFIED reference on several image pairs -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_planner_golden [n_pairs]     # writes tests/golden/planner_pairs.npz / planner_pairs.json

Runs utils.beam_search.beam_search of /root/reference on CPU (import shims of oracle/ref_shims.py) with the settings
of its own driver (preprocess/gen_greedy_seqs_FiveK.py:37-43: beam 3, operations [0,1,2,3,5,6], err 1e-2, L1,
Nelder-Mead, max_step = 6) on seeded pairs whose target is a planted chain of 2-3 operators, and records every beam's
(operator, parameters, distance) sequence.  Minutes of CPU time per pair (a 24-parameter fit is 4 800 evaluations)."""
import json
import os
import sys
import time

import numpy as np
import torch

from . import ops as O
from . import ref_shims
from .make_golden import sample_params

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
GLOBAL_OPS = [0, 1, 2, 3, 5, 6]
PLANTED = [[0, 1], [2, 6], [5, 0], [1, 2, 6], [6, 0], [0, 2], [1, 5], [2, 0, 1]]


def main():
    if not ref_shims.available():
        sys.exit('reference tree not present')
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    R = ref_shims.load()
    opt = R.options()
    torch.manual_seed(10)
    ex = R.executor.Executor(opt)
    H = W = 32
    I0s, Igts, recs = [], [], []
    for i in range(n):
        g = torch.Generator().manual_seed(10 + 3000 + 7 * i)
        I0 = torch.rand(1, 3, H, W, generator=g) * 0.8 + 0.1
        Igt = I0
        with torch.no_grad():
            for op in PLANTED[i % len(PLANTED)]:
                Igt = ex.execute(Igt, op, None, specified_param=sample_params(op, 1, g))[0]
        t0 = time.time()
        actions, Is = R.beam_search.beam_search(I0, Igt, None, ex, None, 3, GLOBAL_OPS, O.ACTION_NAMES, len(GLOBAL_OPS), 1e-2,
                                                'L1', 'Nelder-Mead', replace=False)
        rec = {'planted': PLANTED[i % len(PLANTED)], 'init_dist': R.beam_search.get_dist(I0, Igt, 'L1').item(),
               'actions': [[[a[0], [float(v) for v in a[1]], float(a[2])] for a in seq] for seq in actions],
               'seconds': time.time() - t0}
        print(i, rec['planted'], [[a[0] for a in seq] for seq in rec['actions']], '%.0f s' % rec['seconds'], flush=True)
        I0s.append(I0.numpy()); Igts.append(Igt.numpy()); recs.append(rec)
        np.savez_compressed(os.path.join(OUT, 'planner_pairs.npz'), I0=np.concatenate(I0s), Igt=np.concatenate(Igts))
        with open(os.path.join(OUT, 'planner_pairs.json'), 'w') as f:
            json.dump({'settings': {'beam': 3, 'operations': GLOBAL_OPS, 'max_step': 6, 'err': 1e-2}, 'pairs': recs}, f, indent=1)


if __name__ == '__main__':
    main()
