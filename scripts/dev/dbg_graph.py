import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import t2onet_b200 as T
import t2onet_b200.functional as TF
dev = torch.device('cuda:0')
B, H, W = 8, 64, 64
img = torch.rand(B, 3, H, W, device=dev); tgt = torch.rand_like(img)
ops = torch.tensor([0, 1, 2, 3, 5, 6, 0, 1], device=dev)
params = torch.rand(B, 24, device=dev).requires_grad_()
exe = T.Executor(T.default_options()).to(dev)
feat = torch.randn(B, 512, device=dev)

def try_capture(name, fn):
    try:
        s = torch.cuda.Stream(dev); s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(3): fn()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, 'OK')
    except Exception as e:
        print(name, 'FAILED', repr(e)[:200])
        torch.cuda.synchronize()

try_capture('rows fwd', lambda: TF.execute_rows(img, ops, params.detach()))
def fb():
    params.grad = None
    out = TF.execute_rows(img, ops, params)
    (out - tgt).abs().mean().backward()
try_capture('rows fwd+bwd', fb)
def heads():
    exe.zero_grad(set_to_none=True)
    out, _ = exe.execute_rows(img, ops, None, feat)
    (out - tgt).abs().mean().backward()
try_capture('executor rows fwd+bwd', heads)
try_capture('uniform chain step', lambda: TF.chain_forward_backward(img, [0, 1, 6], [params[:, :1].detach(), params[:, 1:2].detach(), params[:, 2:3].detach()], tgt))

# the bench's episode, step by step
B, H, W, STEPS = 64, 128, 128, 5
gen = torch.Generator().manual_seed(10 + 7000)
img = torch.rand(B, 3, H, W, generator=gen).to(dev); tgt = torch.rand(B, 3, H, W, generator=gen).to(dev)
feats = [torch.randn(B, 512, generator=gen).to(dev) for _ in range(STEPS)]
CH = [0, 1, 2, 3, 5, 6]
ops_dev = [torch.tensor([CH[int(v)] for v in torch.randint(0, 6, (B,), generator=gen)]).to(dev) for _ in range(STEPS)]
def episode(n):
    exe.zero_grad(set_to_none=True)
    x = img
    for k in range(n):
        x = exe.execute_rows(x, ops_dev[k], None, feats[k])[0]
    loss = (x - tgt).abs().mean()
    loss.backward()
    return loss
for n in (1, 2, 5):
    try_capture('episode %d steps' % n, lambda: episode(n))

# with the grouped (reference-style) loop run first, as bench.py does
ops_cpu = [o.cpu() for o in ops_dev]
def grouped_step(x, ops, ctx):
    unqs = torch.unique(ops)
    group_inds = [torch.nonzero(ops == u).squeeze(1) for u in unqs]
    rev = torch.argsort(torch.cat(group_inds)).to(dev)
    outs = []
    for j, inds in enumerate(group_inds):
        inds = inds.to(dev)
        out_g, _ = exe.execute(x.index_select(0, inds), int(unqs[j]), None, ctx.index_select(0, inds), has_noise=False)
        outs.append(out_g)
    return torch.cat(outs).index_select(0, rev)
def episode_g():
    exe.zero_grad(set_to_none=True)
    x = img
    for k in range(STEPS):
        x = grouped_step(x, ops_cpu[k], feats[k])
    loss = (x - tgt).abs().mean(); loss.backward(); return loss
for _ in range(3): episode_g()
torch.cuda.synchronize()
try_capture('episode 5 steps after grouped loop', lambda: episode(5))
import bench as BM
img2, tgt2, prm2 = BM.make_batch(64, 128, 128, 2010, dev)
fs = TF.FusedStep(BM.CHAIN, 64, 128, 128, dev, reuse_outputs=True)
packed = torch.cat(prm2, 1).contiguous()
g0 = torch.cuda.CUDAGraph()
fs(img2, packed, tgt2); torch.cuda.synchronize()
with torch.cuda.graph(g0):
    for _ in range(5): fs(img2, packed, tgt2)
g0.replay(); torch.cuda.synchronize()
try_capture('episode 5 steps after a FusedStep graph', lambda: episode(5))
