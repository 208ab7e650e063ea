"""Golden vectors of the operator classes WITHOUT an Executor slot beyond exposure / white balance -- BNWOperator,
BlurOperator, HueOperator (models/operators.py:298, 373, 414) -- recorded from the UNMODIFIED reference on the inputs
of tests/golden/single_ops.npz.  TEST INFRASTRUCTURE ONLY.     python -m oracle.make_golden_ext

HueOperator.process broadcasts its (B, 1) parameter with `param.expand_as(value)`, which only works for a batch of
one: it is recorded row by row."""
import math
import os
import sys

import numpy as np
import torch

from . import ops as O
from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def sample_params_ext(op, B, g, wide=False):
    u = torch.rand(B, 1, generator=g)
    if op == O.OP_HUE:
        return (u * 2 * math.pi) if not wide else (u * 21 - 7)
    return u if not wide else (u * 2 - 0.5)


def main():
    if not ref_shims.available():
        sys.exit('reference tree not present')
    R = ref_shims.load()
    opt = R.options()
    torch.manual_seed(10)
    cls = {O.OP_BNW: R.operators.BNWOperator(opt), O.OP_BLUR: R.operators.BlurOperator(opt), O.OP_HUE: R.operators.HueOperator(opt)}
    base = np.load(os.path.join(OUT, 'single_ops.npz'))
    img, target, wgt = (torch.from_numpy(base[k]) for k in ('img', 'target', 'wgt'))
    masks = {'none': None, 'm1': torch.from_numpy(base['mask1']), 'm3': torch.from_numpy(base['mask3'])}
    B = img.shape[0]
    g = torch.Generator().manual_seed(10 + 1500)
    rec = {}
    for op in (O.OP_BNW, O.OP_BLUR, O.OP_HUE):
        for tag, wide in (('n', False), ('w', True)):
            p0 = sample_params_ext(op, B, g, wide)
            for mname, mask in masks.items():
                if wide and mname != 'none':
                    continue
                x = img.clone().requires_grad_()
                p = p0.clone().requires_grad_()
                if op == O.OP_HUE:           # batch of one at a time (see the module docstring)
                    out = torch.cat([cls[op].execute(x[b:b + 1], mask=None if mask is None else mask[b:b + 1],
                                                     specified_param=p[b:b + 1]) for b in range(B)], 0)
                else:
                    out = cls[op].execute(x, mask=mask, specified_param=p)
                l1 = (out - target).abs().flatten(1).sum(1)
                (out * wgt).sum().backward()
                key = 'op%d_%s_%s' % (op, tag, mname)
                rec[key + '_param'] = p0.numpy()
                rec[key + '_out'] = out.detach().numpy()
                rec[key + '_l1sum'] = l1.detach().numpy()
                rec[key + '_gimg'] = x.grad.numpy()
                rec[key + '_gparam'] = (p.grad if p.grad is not None else torch.zeros_like(p0)).numpy()
    np.savez_compressed(os.path.join(OUT, 'single_ops_ext.npz'), **rec)
    print('wrote', len(rec), 'arrays')


if __name__ == '__main__':
    main()
