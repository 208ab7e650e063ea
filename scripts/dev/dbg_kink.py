"""Diagnose a parameter-gradient mismatch of the specialised step kernels: which pixels' image gradients differ
between the specialised / generic kernels and the oracle (kink pixels), and by how much the parameter gradients move."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, numpy as np
import t2onet_b200.functional as TF
from oracle import ops as O
from parity_util import *
ops = [0, 1, 2, 3, 5, 6]
B, H, W = 3, 128, 128
g = torch.Generator().manual_seed(79 + H + len(ops))
img = torch.rand(B, 3, H, W, generator=g)
params = [sample_params(op, B, g) for op in ops]
with torch.no_grad():
    target = O.chain(img, ops, [sample_params(op, B, g) for op in ops])
out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, ops, params, target)
res = {}
for mode in ('0', '1'):
    os.environ['T2O_NO_SPECIALIZED'] = mode
    out, l1, grads, gimg = TF.chain_forward_backward(img.cuda(), ops, [p.cuda() for p in params], target.cuda(), want_grad_img=True)
    res[mode] = (out.cpu(), l1.cpu(), [x.cpu() for x in grads], gimg.cpu())
    print('mode', mode, 'numel', img.numel(), '1/numel %.3e' % (1.0 / img.numel()))
    for k, op in enumerate(ops):
        d = (res[mode][2][k] - gp_o[k]).abs()
        print('  op', op, 'max abs diff %.3e' % d.max().item(), 'rel %.3e' % (d.max().item() / gp_o[k].abs().max().item()),
              'at', np.unravel_index(d.argmax().item(), d.shape))
    dg = (res[mode][3] - gi_o).abs()
    idx = (dg > 1e-9 + 1e-4 * gi_o.abs().max()).nonzero()
    print('  grad_img mismatches:', idx.shape[0], 'scale', gi_o.abs().max().item())
    for i in idx[:12].tolist():
        b, c, y, x = i
        print('   px', i, 'ours %.4e oracle %.4e' % (res[mode][3][b, c, y, x].item(), gi_o[b, c, y, x].item()),
              'img', img[b, :, y, x].tolist())
a, b_ = res['0'], res['1']
print('spec vs generic: out %.2e' % (a[0] - b_[0]).abs().max().item(), 'gimg %.2e' % (a[3] - b_[3]).abs().max().item())
for k in range(len(ops)):
    print('  gp', ops[k], '%.3e' % (a[2][k] - b_[2][k]).abs().max().item())
