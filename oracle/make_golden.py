"""Record golden vectors from the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Run in the authoring container (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz, *.json

The reference ships no fixtures for this path, so these are
outputs of the reference's own ``Executor`` / ``Operator.execute`` /
``utils.beam_search`` run on CPU under the import shims of ``oracle/ref_shims.py``
(kornia = ``oracle/hsv.py``: that third-party boundary stays unpinned).
"""
import json
import os
import sys

import numpy as np
import torch

from . import ops as O
from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
GLOBAL_OPS = [0, 1, 2, 3, 5, 6]


def sample_params(op, B, g, wide=False):
    """Parameter distributions of SURVEY.md section 8(d)."""
    n = O.num_params(op)
    u = torch.rand(B, n, generator=g)
    if op == O.OP_BRIGHTNESS:
        return (u * 0.6 - 0.3) if not wide else (u * 4 - 2)
    if op == O.OP_CONTRAST:
        return (u - 0.5) if not wide else (u * 2 - 1)
    if op == O.OP_SATURATION:
        return (u - 0.2) if not wide else (u * 3 - 1.5)
    if op == O.OP_COLOR:
        return 0.9 + 0.2 * u
    if op == O.OP_TONE:
        return 0.5 + 1.5 * u
    if op == O.OP_SHARPNESS:
        return u * 1.5
    if op == O.OP_EXPOSURE:
        return u * 2 - 1
    if op == O.OP_WHITEBALANCE:
        return 0.4 + 1.4 * u
    return u


def adversarial_image(H, W):
    """black / white / gray / two-channel ties / exact curve knots / clamped values."""
    img = torch.zeros(1, 3, H, W)
    vals = [(0, 0, 0), (1, 1, 1), (.5, .5, .5), (.25, .25, .25), (1, 1, .3), (.2, .7, .7), (0, 0, .4), (.6, 0, 0),
            (.125, .375, .875), (.25, .5, .75), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (.9, .1, .1), (.3, .3, .31)]
    k = 0
    for y in range(H):
        for x in range(W):
            img[0, :, y, x] = torch.tensor(vals[k % len(vals)])
            k += 1
    return img


def main():
    if not ref_shims.available():
        sys.exit('reference tree not present; golden vectors can only be regenerated in the authoring container')
    os.makedirs(OUT, exist_ok=True)
    R = ref_shims.load()
    opt = R.options()
    torch.manual_seed(10)
    ex = R.executor.Executor(opt)
    extra = {O.OP_EXPOSURE: R.operators.ExposureOperator(opt), O.OP_WHITEBALANCE: R.operators.ImprovedWhiteBalanceOperator(opt)}

    def ref_execute(img, op, param, mask):
        if op in extra:
            return extra[op].execute(img, mask=mask, specified_param=param)
        return ex.execute(img, op, mask, specified_param=param)[0]

    # ------------------------------------------------------------ single operators
    B, H, W = 3, 12, 20
    g = torch.Generator().manual_seed(10 + 1000)
    img = torch.rand(B, 3, H, W, generator=g)
    img[2:3] = adversarial_image(H, W)
    target = torch.rand(B, 3, H, W, generator=g)
    wgt = torch.randn(B, 3, H, W, generator=g)
    mask1 = (torch.rand(B, 1, H, W, generator=g) > 0.4).float()
    mask3 = torch.rand(B, 3, H, W, generator=g)
    rec = {'img': img.numpy(), 'target': target.numpy(), 'wgt': wgt.numpy(), 'mask1': mask1.numpy(), 'mask3': mask3.numpy()}
    for op in [0, 1, 2, 3, 5, 6, 7, 8, 9]:
        for tag, wide in (('n', False), ('w', True)):
            p0 = sample_params(op, B, g, wide)
            for mname, mask in (('none', None), ('m1', mask1), ('m3', mask3)):
                if wide and mname != 'none':
                    continue
                x = img.clone().requires_grad_()
                p = p0.clone().requires_grad_()
                out = ref_execute(x, op, p, mask)
                l1 = (out - target).abs().flatten(1).sum(1)          # per-image L1 sums
                loss = (out * wgt).sum()
                loss.backward()
                key = 'op%d_%s_%s' % (op, tag, mname)
                rec[key + '_param'] = p0.numpy()
                rec[key + '_out'] = out.detach().numpy()
                rec[key + '_l1sum'] = l1.detach().numpy()
                rec[key + '_gimg'] = x.grad.numpy()
                rec[key + '_gparam'] = (p.grad if p.grad is not None else torch.zeros_like(p0)).numpy()
    np.savez_compressed(os.path.join(OUT, 'single_ops.npz'), **rec)

    # ------------------------------------------------------------ chains (fwd + L1 + grads)
    B, H, W = 2, 16, 24
    g = torch.Generator().manual_seed(10 + 2000)
    img = torch.rand(B, 3, H, W, generator=g)
    rec = {'img': img.numpy()}
    chains = {'c6': [0, 1, 2, 3, 5, 6], 'c6r': [6, 5, 3, 2, 1, 0], 'c3': [1, 6, 5], 'c2': [2, 0]}
    for name, ops in chains.items():
        params_true = [sample_params(op, B, g) for op in ops]
        with torch.no_grad():
            target = img
            for op, p in zip(ops, params_true):
                target = ref_execute(target, op, p, None)
        params = [sample_params(op, B, g).requires_grad_() for op in ops]
        x = img.clone().requires_grad_()
        cur = x
        inter = []
        for op, p in zip(ops, params):
            cur = ref_execute(cur, op, p, None)
            inter.append(cur)
        l1 = (cur - target).abs().mean()          # train_seq2seqL1.py:85
        l1.backward()
        rec[name + '_ops'] = np.array(ops)
        rec[name + '_target'] = target.numpy()
        rec[name + '_out'] = cur.detach().numpy()
        rec[name + '_l1mean'] = np.array(l1.item(), dtype=np.float32)
        rec[name + '_l1dist'] = np.array(R.beam_search.get_dist(cur.detach(), target, 'L1').item(), dtype=np.float32)
        rec[name + '_gimg'] = x.grad.numpy()
        for k, p in enumerate(params):
            rec['%s_param%d' % (name, k)] = p.detach().numpy()
            rec['%s_gparam%d' % (name, k)] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'chains.npz'), **rec)

    # ------------------------------------------------------------ planner transcripts
    H = W = 16
    g = torch.Generator().manual_seed(10 + 3000)
    I0 = torch.rand(1, 3, H, W, generator=g) * 0.8 + 0.1
    with torch.no_grad():
        Igt = ref_execute(I0, 0, torch.tensor([[0.25]]), None)
        Igt = ref_execute(Igt, 1, torch.tensor([[0.4]]), None)
        Igt = ref_execute(Igt, 5, sample_params(5, 1, g), None)
    names = O.ACTION_NAMES
    transcripts = {}
    # per-(state, op) Nelder-Mead fits
    fits = {}
    for op in GLOBAL_OPS:
        calls = [0]
        orig = ex.execute

        def counting(*a, **k):
            calls[0] += 1
            return orig(*a, **k)
        ex.execute = counting
        param, ok = R.beam_search.get_param(I0, Igt, None, op, ex, None, 'L1', 'Nelder-Mead')
        ex.execute = orig
        out = R.beam_search.execute(I0, op, param, ex)
        fits[str(op)] = {'param': param[0].tolist(), 'success': bool(ok), 'nfev': calls[0],
                         'dist': R.beam_search.get_dist(out, Igt, 'L1').item()}
    transcripts['nm_fits'] = fits
    for label, fn, kw in (
            ('beam2', R.beam_search.beam_search, dict(discriminator=None)),
            ('fixed', R.beam_search_fixed_order.beam_search, {}),):
        beam = 2 if label == 'beam2' else 1
        args = [I0, Igt, None, ex]
        if 'discriminator' in kw:
            args.append(None)
        args += [beam, GLOBAL_OPS, names, 3, 1e-2, 'L1', 'Nelder-Mead']
        actions, Is = fn(*args, replace=False)
        transcripts[label] = {'actions': [[[a[0], a[1], a[2]] for a in seq] for seq in actions]}
    transcripts['init_dist'] = R.beam_search.get_dist(I0, Igt, 'L1').item()
    np.savez_compressed(os.path.join(OUT, 'planner_pair.npz'), I0=I0.numpy(), Igt=Igt.numpy())
    with open(os.path.join(OUT, 'planner_transcripts.json'), 'w') as f:
        json.dump(transcripts, f, indent=1)
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
