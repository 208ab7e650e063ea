import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import torch, numpy as np
import t2onet_b200.functional as TF
from oracle import ops as O
from parity_util import *
ops=[0,1,2,3,5,6]
for shape in [(2,40,64),(1,96,260),(3,128,128)]:
    B,H,W=shape
    g = torch.Generator().manual_seed(79 + H + len(ops))
    img = torch.rand(B, 3, H, W, generator=g)
    params = [sample_params(op, B, g) for op in ops]
    with torch.no_grad():
        target = O.chain(img, ops, [sample_params(op, B, g) for op in ops])
    out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, ops, params, target)
    for mode in ('0','1'):
        os.environ['T2O_NO_SPECIALIZED']=mode
        out, l1, grads, gimg = TF.chain_forward_backward(img.cuda(), ops, [p.cuda() for p in params], target.cuda(), want_grad_img=True)
        d=(out.cpu()-out_o).abs()
        i=d.argmax().item()
        print(shape, mode, 'max', d.max().item(), 'ours', out.cpu().flatten()[i].item(), 'oracle', out_o.flatten()[i].item(), 'p_sharp', params[5].flatten().tolist())
