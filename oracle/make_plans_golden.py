"""Record what the UNMODIFIED reference's FiveKAct.get_act (datasets/FiveKdataset.py:86-113) and analyze_traj (:54-64)
return for synthetic planner records -- TEST INFRASTRUCTURE ONLY.   python -m oracle.make_plans_golden"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    if not ref_shims.available():
        sys.exit('reference tree not present')
    ref_shims.load()
    import cv2
    sys.modules.setdefault('cv2', cv2)
    from datasets import FiveKdataset as FD
    rng = np.random.default_rng(10)
    names = ['brightness', 'contrast', 'saturation', 'color', 'tone', 'sharpness']
    pn = {'brightness': 1, 'contrast': 1, 'saturation': 1, 'color': 24, 'tone': 8, 'sharpness': 1}
    cases = []
    tmp = tempfile.mkdtemp()
    for i in range(12):
        k = int(rng.integers(1, 7))
        ops = list(rng.permutation(names)[:k])
        init = float(rng.uniform(0.05, 0.3))
        d, seq = init, []
        for j, op in enumerate(ops):
            step = rng.choice([0.3, 0.05, 0.009, 0.0, -0.01], p=[0.45, 0.35, 0.08, 0.06, 0.06]) * init      # improvements above / below the 1 % rule
            d = max(d - step, 1e-4)
            prm = rng.uniform(0.5, 2.0, pn[op]) if pn[op] > 1 else rng.choice([rng.uniform(-1, 1), 7.5], size=1, p=[0.85, 0.15])
            seq.append([op, [float(v) for v in prm], float(d)])
        rec = {'request': 'req %d' % i, 'init distance': init, 'operation sequence': [seq, seq[:1]]}
        item_dir = os.path.join(tmp, 'train%d' % i)
        os.makedirs(item_dir)
        with open(os.path.join(item_dir, '%05d.json' % i), 'w') as f:
            json.dump(rec, f)
        trunc = min(FD.analyze_traj([init] + [a[2] for a in seq]), 5)
        for j in range(trunc):
            cv2.imwrite(os.path.join(item_dir, 'edit%d.jpg' % j), np.zeros((8, 8, 3), np.uint8))
        fake = types.SimpleNamespace(act_dir=tmp, phase='train', op_max_len=5, train_img_size=8,
                                     actions=['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white'],
                                     act2pn={'brightness': 1, 'contrast': 1, 'saturation': 1, 'color': 24, 'inpaint': 0, 'tone': 8,
                                             'sharpness': 1, 'white': 0})
        op_seq, params, imgs = FD.FiveKAct.get_act(fake, i)
        cases.append({'record': rec, 'op_seq': [int(v) for v in op_seq], 'params': params.tolist(), 'trunc_len': int(trunc)})
    with open(os.path.join(OUT, 'plans.json'), 'w') as f:
        json.dump(cases, f)
    print('wrote', len(cases), 'cases;', [c['trunc_len'] for c in cases])


if __name__ == '__main__':
    main()
