"""Record FULL planner transcripts of the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_planner_golden_full c3   [n_pairs]   # 3x128x128, beam 8 (BASELINE config 3's shape)
    python -m oracle.make_planner_golden_full c5   [n_pairs]   # 3x256x256, beam 8 (GIER-shaped, config 5)
    python -m oracle.make_planner_golden_full eps  [n_pairs]   # utils/beam_search_eps_greedy.py, random.seed(0), 32x32
    python -m oracle.make_planner_golden_full c3a | c5a [n]    # the same pairs with the L1 summed in float64 (see below)

writes tests/golden/planner_full_<mode>.npz (uint8 image pairs) and planner_full_<mode>.json.

Runs utils.beam_search.beam_search of /root/reference on CPU (import shims of oracle/ref_shims.py) with the settings of
BASELINE config 3 (preprocess/gen_greedy_seqs_FiveK.py:37-43 with beam 8: operations [0,1,2,3,5,6], err 1e-2, L1,
Nelder-Mead, max_step 6).  Unlike make_planner_golden.py it records, per beam step, EVERY candidate the reference
evaluated -- (parent beam index, parent operator sequence, operator, fitted parameters, distance, kept or not) -- and
the distance array / order of its np.argsort, so that a test can show whether a candidate displaced in another
implementation's run was tied within tolerance in the reference's own run.  The reference's code is not modified: the
module-level names `get_param`, `get_dist`, `minimize` and `np` that beam_search looks up are wrapped by recording
pass-throughs.  For the 1-parameter fits the whole Nelder-Mead evaluation history (x, f) is kept as well.

The `a` modes (c3a, c5a: json only, the images are those of c3 / c5) change ONE thing, and say so in their settings
('l1_sum': 'float64'): get_dist's `(x1 - x2).norm(1)` (utils/beam_search.py:173) is summed in float64 and rounded to the
tensors' dtype.  Why: on the CPU torch sums an fp32 norm(1) in a few long fp32 accumulator chains, which at 49 152 /
196 608 elements carries 1e-5 / 1e-4 of accumulation noise (measured against float64; it depends on the thread count and the
vector width) -- more than the change of the L1 across Nelder-Mead's first steps of 2.5e-4, so the unmodified reference's
1-parameter fits stop inside that noise after 6-14 evaluations on this host.  On the device the reference is written
for (utils/beam_search.py:29, CUDA) norm(1) is a tree reduction that is accurate to an ulp or two; the `a` transcripts
are what the reference's planner does with such a sum.  Both kinds are committed and tested.

Inputs are 8-bit images (x / 255, as utils/visual_utils.py:61-70 produces them): a smooth random colour field plus
noise, the target a planted chain of 2-4 operators re-quantised to 8 bits, so the fixture stores uint8.
One process per pair (torch threads = 1); a pair takes minutes of CPU time (a 24-parameter fit is 4 800 evaluations)."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
GLOBAL_OPS = [0, 1, 2, 3, 5, 6]
PLANTED = [[0, 1], [2, 6], [5, 0], [1, 2, 6], [6, 0], [0, 2], [1, 5], [2, 0, 1], [3, 0], [5, 2, 6], [0, 3, 1], [1, 6],
           [2, 5], [3, 6, 0], [0, 1, 2, 5], [6, 2, 1]]
MODES = {'c3': dict(H=128, W=128, beam=8, seed=4000, variant='default'),
         'c5': dict(H=256, W=256, beam=8, seed=5000, variant='default'),
         'c3a': dict(H=128, W=128, beam=8, seed=4000, variant='default', accurate_l1=True, images='c3'),
         'c5a': dict(H=256, W=256, beam=8, seed=5000, variant='default', accurate_l1=True, images='c5'),
         'eps': dict(H=32, W=32, beam=8, seed=6000, variant='eps_greedy')}


def make_pair(i, H, W, seed, ex):
    """Seeded 8-bit pair: smooth colour field + noise, target = planted chain, both quantised to k/255."""
    from .make_golden import sample_params
    g = torch.Generator().manual_seed(10 + seed + 7 * i)
    coarse = torch.rand(1, 3, 8, 8, generator=g)
    base = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False)
    I0 = (base * 0.7 + 0.15 + (torch.rand(1, 3, H, W, generator=g) - 0.5) * 0.2).clamp(0.02, 0.98)
    I0 = torch.round(I0 * 255).to(torch.uint8)
    x = I0.float() / 255
    with torch.no_grad():
        for op in PLANTED[i % len(PLANTED)]:
            x = ex.execute(x, op, None, specified_param=sample_params(op, 1, g))[0]
    Igt = torch.round(x.clamp(0, 1) * 255).to(torch.uint8)
    return I0, Igt


class _NumpyProxy:
    """Stands in for the module global `np` of the reference planner: forwards everything, records argsort."""

    def __init__(self, rec):
        self._rec = rec

    def __getattr__(self, name):
        return getattr(np, name)

    def argsort(self, a, *args, **kw):
        order = np.argsort(a, *args, **kw)
        self._rec['sort'].append({'dists': [float(v) for v in a], 'order': [int(v) for v in order]})
        return order


def run_pair(job):
    mode, i, eps = job
    torch.set_num_threads(1)
    from . import ops as O
    from . import ref_shims
    cfg = MODES[mode]
    R = ref_shims.load()
    torch.manual_seed(10)
    ex = R.executor.Executor(R.options())
    I0u, Igtu = make_pair(i, cfg['H'], cfg['W'], cfg['seed'], ex)
    I0, Igt = I0u.float() / 255, Igtu.float() / 255
    mod = R.beam_search_eps_greedy if cfg['variant'] == 'eps_greedy' else R.beam_search
    rec = {'sort': [], 'cands': []}
    state = {'in_fit': False, 'op': None, 'parent': None, 'param': None, 'nfev': 0}
    orig_get_param, orig_get_dist, orig_np, orig_minimize = mod.get_param, mod.get_dist, mod.np, mod.minimize

    def minimize(func, x0, **kw):
        """scipy.optimize.minimize as the reference calls it, with every evaluation (x, f) of the fit recorded"""
        hist = []

        def f2(x):
            v = func(x)
            hist.append(([float(t) for t in np.atleast_1d(x)], float(v)))
            return v
        res = orig_minimize(f2, x0, **kw)
        state['hist'] = hist
        return res

    def get_param(I, I_gt, txt, operation, *a, **k):
        state['in_fit'], state['op'], state['nfev'], state['hist'] = True, operation, 0, None
        state['parent'] = I
        try:
            param, ok = orig_get_param(I, I_gt, txt, operation, *a, **k)
        finally:
            state['in_fit'] = False
        state['param'] = param
        return param, ok

    def get_dist(x1, x2, dist_type):
        if cfg.get('accurate_l1'):
            assert dist_type == 'L1'
            d = ((x1 - x2).double().norm(1) / x1.numel()).to(torch.result_type(x1, x2))
        else:
            d = orig_get_dist(x1, x2, dist_type)
        if state['in_fit']:
            state['nfev'] += 1
        else:
            rec['cands'].append({'step': len(rec['sort']), 'op': int(state['op']),
                                 'param': [float(v) for v in state['param'][0].tolist()], 'dist': float(d.item()),
                                 'nfev': state['nfev'], 'parent_id': id(state['parent']),
                                 # the whole Nelder-Mead evaluation history of the 1-parameter fits (x, f): lets a test
                                 # show whether a fit that ends elsewhere was decided by two evaluations tied within the
                                 # L1's own rounding noise (the 8- / 24-parameter histories would be ~0.5 MB per fit)
                                 'hist': [[h[0][0], h[1]] for h in state['hist']] if len(state['param'][0]) == 1 else None})
        return d

    mod.get_param, mod.get_dist, mod.np, mod.minimize = get_param, get_dist, _NumpyProxy(rec), minimize
    t0 = time.time()
    try:
        if cfg['variant'] == 'eps_greedy':
            mod.random.seed(0)                                   # utils/beam_search_eps_greedy.py:24
            actions, Is = mod.beam_search(I0, Igt, None, ex, None, cfg['beam'], GLOBAL_OPS, O.ACTION_NAMES, len(GLOBAL_OPS),
                                          1e-2, 'L1', 'Nelder-Mead', eps=eps, replace=False)
        else:
            actions, Is = mod.beam_search(I0, Igt, None, ex, None, cfg['beam'], GLOBAL_OPS, O.ACTION_NAMES, len(GLOBAL_OPS),
                                          1e-2, 'L1', 'Nelder-Mead', replace=False)
        init_dist = orig_get_dist(I0, Igt, 'L1').item()
    finally:
        mod.get_param, mod.get_dist, mod.np, mod.minimize = orig_get_param, orig_get_dist, orig_np, orig_minimize
    # parent tensors -> the beam index they had in that step's I_buff (first-seen order within the step)
    steps = []
    for s in range(len(rec['sort'])):
        cs = [c for c in rec['cands'] if c['step'] == s]
        seen = []
        for c in cs:
            if c['parent_id'] not in seen:
                seen.append(c['parent_id'])
            c['parent'] = seen.index(c['parent_id'])
        steps.append({'candidates': [{k: c[k] for k in ('parent', 'op', 'param', 'dist', 'nfev', 'hist')} for c in cs],
                      'sort_dists': rec['sort'][s]['dists'], 'sort_order': rec['sort'][s]['order']})
    out = {'index': i, 'planted': PLANTED[i % len(PLANTED)], 'init_dist': init_dist, 'eps': eps,
           'actions': [[[a[0], [float(v) for v in a[1]], float(a[2])] for a in seq] for seq in actions],
           'steps': steps, 'seconds': time.time() - t0}
    print(mode, i, out['planted'], [[a[0] for a in seq] for seq in out['actions']][:2], '%d steps' % len(steps),
          '%.0f s' % out['seconds'], flush=True)
    return out, I0u.numpy(), Igtu.numpy()


def main():
    from . import ref_shims
    if not ref_shims.available():
        sys.exit('reference tree not present')
    mode = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cfg = MODES[mode]
    if mode == 'eps':
        # eps = 0.05 (the reference's default: the first draw of random.seed(0) is 0.844 -> greedy branch) and
        # eps = 0.9 (the random.choices branch), alternating
        jobs = [(mode, i, 0.05 if i % 2 == 0 else 0.9) for i in range(n)]
    else:
        jobs = [(mode, i, None) for i in range(n)]
    with mp.get_context('spawn').Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(run_pair, jobs, chunksize=1)
    if cfg.get('images'):
        d = np.load(os.path.join(OUT, 'planner_full_%s.npz' % cfg['images']))      # same seeds -> the same pairs
        assert np.array_equal(d['I0'][:n], np.concatenate([r[1] for r in res])) and np.array_equal(d['Igt'][:n], np.concatenate([r[2] for r in res]))
    else:
        np.savez_compressed(os.path.join(OUT, 'planner_full_%s.npz' % mode), I0=np.concatenate([r[1] for r in res]),
                            Igt=np.concatenate([r[2] for r in res]))
    with open(os.path.join(OUT, 'planner_full_%s.json' % mode), 'w') as f:
        json.dump({'settings': {'beam': cfg['beam'], 'operations': GLOBAL_OPS, 'max_step': len(GLOBAL_OPS), 'err': 1e-2,
                                'variant': cfg['variant'], 'shape': [3, cfg['H'], cfg['W']],
                                'l1_sum': 'float64' if cfg.get('accurate_l1') else 'torch CPU fp32 norm(1), 1 thread',
                                'images': 'planner_full_%s.npz' % cfg.get('images', mode)},
                   'pairs': [r[0] for r in res]}, f, separators=(',', ':'))


if __name__ == '__main__':
    main()
