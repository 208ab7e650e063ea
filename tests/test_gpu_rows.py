"""Parity of the per-row operator step (t2o_rows_forward / t2o_rows_backward, Executor.execute_rows) against the
CPU oracle: every batch row applies its OWN operator, as the Actor's divide_op_group loop does
(models/actor.py:100-114, 156-170, 245-259).  Tolerances: max-abs 1e-5 on pixels / L1, relative 1e-4 on gradients."""
import numpy as np
import pytest
import torch

from oracle import ops as O
from parity_util import TOL_GRAD, TOL_PIX, max_abs, rel_err, rel_err_kinks, sample_params

pytestmark = pytest.mark.gpu
SLOT = 24


@pytest.fixture(scope='module')
def TF():
    import t2onet_b200.functional as TF
    return TF


def make_rows(row_ops, H, W, seed, mask_ch=0):
    """row_ops: list (B) of lists (K).  Returns img, params (B, K*24) zero-padded, target, wgt, mask."""
    g = torch.Generator().manual_seed(seed)
    B, K = len(row_ops), len(row_ops[0])
    img = torch.rand(B, 3, H, W, generator=g)
    target = torch.rand(B, 3, H, W, generator=g)
    wgt = torch.randn(B, 3, H, W, generator=g)
    params = torch.zeros(B, K * SLOT)
    for b in range(B):
        for k, op in enumerate(row_ops[b]):
            if op >= 0:
                p = sample_params(op, 1, g)
                params[b, k * SLOT:k * SLOT + p.shape[1]] = p[0]
    mask = (torch.rand(B, mask_ch, H, W, generator=g) > 0.4).float() * torch.rand(B, mask_ch, H, W, generator=g) if mask_ch else None
    return img, params, target, wgt, mask


def oracle_rows(img, row_ops, params, mask, loss_fn):
    """Row by row through the oracle; returns out, grad_params (same layout as params), grad_img."""
    x = img.clone().requires_grad_()
    p = params.clone().requires_grad_()
    outs = []
    for b, ops in enumerate(row_ops):
        ps = [p[b:b + 1, k * SLOT:k * SLOT + max(O.num_params(op), 1)] if op >= 0 else None for k, op in enumerate(ops)]
        m = None if mask is None else mask[b:b + 1]
        outs.append(O.chain(x[b:b + 1], ops, ps, m))
    out = torch.cat(outs)
    loss_fn(out).backward()
    return out.detach(), p.grad, x.grad


MIXED = [[0], [1], [2], [3], [5], [6], [7], [-1], [6], [0], [8], [9], [5], [3], [10], [11], [12], [11]]


@pytest.mark.parametrize('shape', [(32, 48), (37, 53), (128, 128), (9, 6)])
@pytest.mark.parametrize('mask_ch', [0, 1, 3])
@pytest.mark.parametrize('where', ['host', 'device'])
def test_rows_single_step(TF, shape, mask_ch, where):
    H, W = shape
    img, params, target, wgt, mask = make_rows(MIXED, H, W, 10 + H * W + mask_ch, mask_ch)
    out_o, gp_o, gi_o = oracle_rows(img, MIXED, params, mask, lambda o: (o * wgt).sum())
    x = img.cuda().requires_grad_()
    p = params.cuda().requires_grad_()
    ops = [r[0] for r in MIXED]
    row_ops = torch.tensor(ops).cuda() if where == 'device' else ops
    out = TF.execute_rows(x, row_ops, p, None if mask is None else mask.cuda())
    (out * wgt.cuda()).sum().backward()
    assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
    # the loss is a signed sum of 3*H*W terms of magnitude ~1: a row whose gradient cancels to ~0 carries fp32
    # summation-order noise of about eps * sqrt(3 H W), which is not a relative error of the kernel
    atol = 1e-7 * (3 * H * W) ** 0.5
    for b, op in enumerate(ops):
        if op in (7, -1):
            assert float(p.grad[b].abs().max()) == 0.0
            continue
        assert rel_err(p.grad[b].cpu(), gp_o[b], atol=atol) <= TOL_GRAD, 'row %d op %d' % (b, op)
    assert rel_err(x.grad.cpu(), gi_o) <= TOL_GRAD
    assert TF.rows_status(x.device) == 0


CHAIN_ROWS = [[0, 1, 6], [6, 5, 3], [2, 6, 0], [3, 5, 1], [-1, -1, -1], [0, -1, 5], [6, -1, -1], [-1, 6, 2], [8, 9, 7], [1, 2, -1]]


@pytest.mark.parametrize('shape', [(40, 64), (67, 131), (128, 128)])
@pytest.mark.parametrize('mask_ch', [0, 1])
def test_rows_chains_host_ops(TF, shape, mask_ch):
    H, W = shape
    img, params, target, wgt, mask = make_rows(CHAIN_ROWS, H, W, 77 + H, mask_ch)
    out_o, gp_o, gi_o = oracle_rows(img, CHAIN_ROWS, params, mask, lambda o: (o - target).abs().mean())
    l1_o = (out_o - target).abs().flatten(1).sum(1)
    mc = None if mask is None else mask.cuda()
    # autograd path: forward launch, then backward launch fed by grad_out
    x = img.cuda().requires_grad_()
    p = params.cuda().requires_grad_()
    out = TF.execute_rows(x, CHAIN_ROWS, p, mc)
    (out - target.cuda()).abs().mean().backward()
    assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
    # one kink pixel (see rel_err_kinks) moves a mean-L1 parameter gradient by up to ~1/numel = 1.3e-5 * |dy/dp|:
    # absolute slack of 5e-6 (row 0 of shape0 has two such pixels, measured 2e-6)
    assert rel_err(p.grad.cpu(), gp_o, atol=5e-6) <= TOL_GRAD
    assert rel_err_kinks(x.grad.cpu(), gi_o) <= TOL_GRAD
    # fused step: forward + L1 + backward in one launch per tiling
    out2, l1, gp, gi = TF.rows_forward_backward(img.cuda(), CHAIN_ROWS, params.cuda(), target.cuda(), mc, want_grad_img=True)
    assert max_abs(out2.cpu(), out_o) <= TOL_PIX
    assert np.allclose(l1.cpu().numpy(), l1_o.numpy(), rtol=3e-6, atol=1e-4)
    for b, ops in enumerate(CHAIN_ROWS):
        assert rel_err(gp[b].cpu(), gp_o[b], atol=5e-6) <= TOL_GRAD, 'row %d ops %s' % (b, ops)
    assert rel_err_kinks(gi.cpu(), gi_o) <= TOL_GRAD


def test_rows_match_uniform_chain(TF):
    """All rows the same chain: the per-row kernels must reproduce the uniform-chain kernels bit for bit."""
    ops = [0, 1, 2, 3, 5, 6]
    B, H, W = 4, 64, 96
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 3, H, W, generator=g).cuda()
    target = torch.rand(B, 3, H, W, generator=g).cuda()
    plist = [sample_params(op, B, g).cuda() for op in ops]
    params = torch.zeros(B, len(ops) * SLOT, device='cuda')
    for k, pk in enumerate(plist):
        params[:, k * SLOT:k * SLOT + pk.shape[1]] = pk
    out_u, l1_u, grads_u, gi_u = TF.chain_forward_backward(img, ops, plist, target, want_grad_img=True)
    out_r, l1_r, gp_r, gi_r = TF.rows_forward_backward(img, [ops] * B, params, target, want_grad_img=True)
    assert torch.equal(out_u, out_r) and torch.equal(l1_u, l1_r) and torch.equal(gi_u, gi_r)
    for k, gk in enumerate(grads_u):
        assert torch.equal(gk, gp_r[:, k * SLOT:k * SLOT + gk.shape[1]])


def test_rows_validation(TF):
    import t2onet_b200._lib as L
    B, H, W = 3, 16, 16
    img = torch.rand(B, 3, H, W).cuda()
    # host-known rows: inpaint and a duplicated operator type (backward) are refused up front
    with pytest.raises(L.T2OError):
        TF.execute_rows(img, [0, 4, 1], torch.zeros(B, SLOT).cuda())
    with pytest.raises(L.T2OError):
        TF.rows_forward_backward(img, [[0, 0], [1, 2], [3, 5]], torch.ones(B, 2 * SLOT).cuda(), img)
    with pytest.raises(L.T2OError):       # a second stencil in one row
        TF.execute_rows(img, [[6, 6], [1, 2], [3, 5]], torch.ones(B, 2 * SLOT).cuda())
    with pytest.raises(L.T2OError):       # device-resident ids need K == 1
        TF.execute_rows(img, torch.tensor([[0, 1], [1, 2], [3, 5]]).cuda(), torch.ones(B, 2 * SLOT).cuda())
    with pytest.raises(L.T2OError):       # parameter table of the wrong width
        TF.execute_rows(img, [0, 1, 2], torch.zeros(B, 8).cuda())
    # device-resident rows: an invalid id is treated as identity and flagged
    assert TF.rows_status(img.device) == 0
    out = TF.execute_rows(img, torch.tensor([0, 42, 4]).cuda(), torch.zeros(B, SLOT).cuda())
    assert TF.rows_status(img.device) == 1 and TF.rows_status(img.device) == 0
    assert torch.equal(out[1:], img[1:])


@pytest.mark.parametrize('where', ['host', 'device'])
@pytest.mark.parametrize('use_mask', [False, True])
def test_executor_execute_rows_vs_grouped_loop(where, use_mask):
    """Executor.execute_rows == the Actor's divide_op_group loop over Executor.execute (models/actor.py:156-170),
    values and gradients down to the FC-head weights."""
    import t2onet_b200 as T
    torch.manual_seed(10)
    ex = T.Executor(T.default_options()).cuda()
    bs, H, W = 16, 32, 32
    g = torch.Generator().manual_seed(3)
    img = torch.rand(bs, 3, H, W, generator=g).cuda()
    feat = torch.randn(bs, 512, generator=g).cuda()
    wgt = torch.randn(bs, 3, H, W, generator=g).cuda()
    mask = (torch.rand(bs, 1, H, W, generator=g) > 0.5).float().cuda() if use_mask else None
    vocab_ops = torch.tensor([3, 4, 5, 6, 8, 9, 10, 2] * 2)          # vocab ids; Executor index = id - 3, <END> = 2
    ops = vocab_ops - 3

    def grouped(img_x, context):
        # the reference loop, restated
        out_gs, par_gs, group_inds = [], [], []
        unqs = torch.unique(ops)
        for unq in unqs:
            group_inds.append(torch.nonzero(ops == unq).squeeze(1))
        rev = torch.argsort(torch.cat(group_inds))
        for j, inds in enumerate(group_inds):
            inds = inds.cuda()
            img_g, ctx_g = img_x.index_select(0, inds), context.index_select(0, inds)
            mask_g = mask.index_select(0, inds) if mask is not None else None
            out_g, par_g = ex.execute(img_g, int(unqs[j]), mask_g, ctx_g, has_noise=False)
            par_gs.append(torch.cat([par_g, torch.zeros(len(inds), 24 - par_g.shape[-1], device=par_g.device)], 1))
            out_gs.append(out_g)
        return torch.cat(out_gs).index_select(0, rev.cuda()), torch.cat(par_gs).index_select(0, rev.cuda())

    x1, f1 = img.clone().requires_grad_(), feat.clone().requires_grad_()
    out1, par1 = grouped(x1, f1)
    ex.zero_grad()
    ((out1 * wgt).sum() + par1.sum()).backward()
    ref_grads = {n: p.grad.clone() for n, p in ex.named_parameters() if p.grad is not None}
    x2, f2 = img.clone().requires_grad_(), feat.clone().requires_grad_()
    ex.zero_grad()
    out2, par2 = ex.execute_rows(x2, ops.cuda() if where == 'device' else ops, mask, f2)
    ((out2 * wgt).sum() + par2.sum()).backward()
    assert max_abs(out2.detach().cpu(), out1.detach().cpu()) <= TOL_PIX
    assert max_abs(par2.detach().cpu(), par1.detach().cpu()) <= 1e-6
    assert rel_err(x2.grad.cpu(), x1.grad.cpu()) <= TOL_GRAD
    assert rel_err(f2.grad.cpu(), f1.grad.cpu()) <= TOL_GRAD
    for n, p in ex.named_parameters():
        if n in ref_grads:
            assert rel_err(p.grad.cpu(), ref_grads[n].cpu(), atol=1e-6) <= TOL_GRAD, n


def test_executor_batched_heads_match_per_operator_heads():
    """execute_rows(batched_heads=True): all FC heads as two batched GEMMs == one pair of small GEMMs per operator, up to the
    GEMM's summation order (values 1e-5, gradients down to the FC weights relative 1e-4)."""
    import t2onet_b200 as T
    torch.manual_seed(10)
    ex = T.Executor(T.default_options()).cuda()
    bs, H, W = 16, 32, 32
    g = torch.Generator().manual_seed(4)
    img = torch.rand(bs, 3, H, W, generator=g).cuda()
    feat = torch.randn(bs, 512, generator=g).cuda()
    wgt = torch.randn(bs, 3, H, W, generator=g).cuda()
    ops = (torch.tensor([3, 4, 5, 6, 8, 9, 10, 2] * 2) - 3).cuda()
    res = []
    for batched in (False, True):
        x, f = img.clone().requires_grad_(), feat.clone().requires_grad_()
        ex.zero_grad()
        out, par = ex.execute_rows(x, ops, None, f, batched_heads=batched)
        ((out * wgt).sum() + par.sum()).backward()
        res.append((out.detach().cpu(), par.detach().cpu(), x.grad.cpu(), f.grad.cpu(),
                    {n: p.grad.clone().cpu() for n, p in ex.named_parameters() if p.grad is not None}))
    a, b = res
    assert max_abs(a[0], b[0]) <= TOL_PIX and max_abs(a[1], b[1]) <= 1e-5
    assert rel_err(b[2], a[2]) <= TOL_GRAD and rel_err(b[3], a[3]) <= TOL_GRAD
    assert set(a[4]) == set(b[4])
    for n in a[4]:
        assert rel_err(b[4][n], a[4][n], atol=1e-6) <= TOL_GRAD, n
