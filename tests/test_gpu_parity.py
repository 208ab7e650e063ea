"""Parity of the sm_100a kernels (through the C-ABI, via the Python binding) against the CPU oracle
and the golden vectors recorded from the reference.  Tolerances are the north star's:
max-abs 1e-5 on edited pixels and L1 (fp32), relative 1e-4 on parameter gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import ops as O
from parity_util import TOL_GRAD, TOL_PIX, kink_slack, max_abs, oracle_chain_with_grads, rel_err, rel_err_kinks, sample_params

pytestmark = pytest.mark.gpu

ALL_OPS = [0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12]       # 10-12: BNW, Blur, Hue (single_ops_ext.npz)


@pytest.fixture(scope='module')
def TF():
    import t2onet_b200.functional as TF
    return TF


@pytest.fixture(scope='module')
def single(golden_dir):
    d = dict(np.load(os.path.join(golden_dir, 'single_ops.npz')))
    d.update(np.load(os.path.join(golden_dir, 'single_ops_ext.npz')))
    return d


@pytest.fixture(scope='module')
def chains(golden_dir):
    return np.load(os.path.join(golden_dir, 'chains.npz'))


def cuda(t):
    return None if t is None else torch.as_tensor(t).cuda()


# ------------------------------------------------------------------ golden vectors from the reference
@pytest.mark.parametrize('op', ALL_OPS)
@pytest.mark.parametrize('variant', ['n_none', 'n_m1', 'n_m3', 'w_none'])
def test_single_op_golden(TF, single, op, variant):
    key = 'op%d_%s' % (op, variant)
    mk = variant.split('_')[1]
    mask = None if mk == 'none' else cuda(single['mask1' if mk == 'm1' else 'mask3'])
    img = cuda(single['img']).requires_grad_()
    p = cuda(single[key + '_param']).requires_grad_()
    out = TF.chain(img, [op], [p], mask)
    (out * cuda(single['wgt'])).sum().backward()
    assert max_abs(out.detach().cpu(), single[key + '_out']) <= TOL_PIX
    l1 = TF.chain_l1(img.detach(), [op], [p.detach()], cuda(single['target']), mask)
    assert np.allclose(l1.cpu().numpy(), single[key + '_l1sum'], rtol=2e-6, atol=TOL_PIX * 10)
    if op != 7:
        assert rel_err(p.grad.cpu(), single[key + '_gparam']) <= TOL_GRAD
    # image gradient: images 0,1 are random (no ties); image 2 is the adversarial tie image, where only
    # the operators without channel max/min routing are required to match (gradient ties)
    nimg = 3 if op not in (0, 2) else 2
    assert rel_err(img.grad[:nimg].cpu(), single[key + '_gimg'][:nimg]) <= TOL_GRAD


@pytest.mark.parametrize('name', ['c6', 'c6r', 'c3', 'c2'])
def test_chain_golden(TF, chains, name):
    ops = [int(v) for v in chains[name + '_ops']]
    img = cuda(chains['img']).requires_grad_()
    params = [cuda(chains['%s_param%d' % (name, k)]).requires_grad_() for k in range(len(ops))]
    target = cuda(chains[name + '_target'])
    out = TF.chain(img, ops, params)
    assert max_abs(out.detach().cpu(), chains[name + '_out']) <= TOL_PIX
    l1 = TF.chain_l1(img, ops, params, target)
    loss = l1.sum() / img.numel()
    assert abs(loss.item() - float(chains[name + '_l1mean'])) <= TOL_PIX
    loss.backward()
    for k, p in enumerate(params):
        assert rel_err(p.grad.cpu(), chains['%s_gparam%d' % (name, k)]) <= TOL_GRAD
    assert rel_err(img.grad.cpu(), chains[name + '_gimg']) <= TOL_GRAD
    # fused forward+backward in one launch gives the same numbers
    out2, l12, grads, gi = TF.chain_forward_backward(img.detach(), ops, [p.detach() for p in params], target,
                                                     want_out=True, want_grad_img=True)
    assert max_abs(out2.cpu(), chains[name + '_out']) <= TOL_PIX
    assert np.allclose(l12.cpu().numpy(), l1.detach().cpu().numpy(), rtol=2e-6)   # other tiling, other summation order
    for k, gk in enumerate(grads):
        assert rel_err(gk.cpu(), chains['%s_gparam%d' % (name, k)]) <= TOL_GRAD
    assert rel_err(gi.cpu(), chains[name + '_gimg']) <= TOL_GRAD


# ------------------------------------------------------------------ seeded inputs against the oracle
SHAPES = [(2, 32, 48), (1, 37, 53), (3, 8, 4), (1, 1, 1), (2, 64, 130), (1, 128, 128), (1, 33, 6)]


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('op', ALL_OPS)
def test_single_op_oracle(TF, op, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(10 + 100 * op + H * W)
    img = torch.rand(B, 3, H, W, generator=g)
    target = torch.rand(B, 3, H, W, generator=g)
    wgt = torch.randn(B, 3, H, W, generator=g)
    p = sample_params(op, B, g)
    out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, [op], [p], target, None, wgt)
    x = img.cuda().requires_grad_()
    pc = p.cuda().requires_grad_()
    out = TF.chain(x, [op], [pc])
    (out * wgt.cuda()).sum().backward()
    assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
    l1 = TF.chain_l1(img.cuda(), [op], [p.cuda()], target.cuda())
    assert np.allclose(l1.cpu().numpy(), l1_o.numpy(), rtol=3e-6, atol=1e-5)
    if op != 7:
        assert rel_err(pc.grad.cpu(), gp_o[0]) <= TOL_GRAD
    assert rel_err(x.grad.cpu(), gi_o) <= TOL_GRAD
    # Operator.process (no blend, no clamp)
    raw = TF.process_raw(img.cuda(), op, p.cuda())
    assert max_abs(raw.cpu(), O.process(op, img, p)) <= TOL_PIX * 4   # unclamped values reach ~13


CHAINS = [[0, 1, 2, 3, 5, 6], [6, 0, 1, 2, 3, 5], [1, 6, 5], [3, 5], [6], [2, 6, 0], [8, 9, 7, 0], [0, 1, 2, 3, 5, 8, 9, 6],
          [6, 1, 6, 5], [5, 5, 3, 3, 1, 1, 0, 0, 2, 2], [0, -1, 5],
          [10, 11, 12], [11], [0, 11, 5, 10], [12, 6, 11, 1], [2, 12, 3, 11]]        # BNW 10, Blur 11, Hue 12


@pytest.mark.parametrize('shape', [(2, 40, 64), (1, 67, 131), (2, 128, 128), (1, 35, 66)])
@pytest.mark.parametrize('ops', CHAINS, ids=lambda c: '-'.join(map(str, c)))
def test_chain_oracle(TF, ops, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(10 + 7 * len(ops) + sum(ops) + H)
    img = torch.rand(B, 3, H, W, generator=g)
    params = [sample_params(op, B, g) if op >= 0 else torch.zeros(B, 1) for op in ops]
    with torch.no_grad():
        target = O.chain(img, ops, [sample_params(op, B, g) if op >= 0 else None for op in ops])
    out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, ops, params, target)
    x = img.cuda().requires_grad_()
    ps = [p.cuda().requires_grad_() for p in params]
    out = TF.chain(x, ops, ps)
    assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
    l1 = TF.chain_l1(x, ops, ps, target.cuda())
    assert np.allclose(l1.detach().cpu().numpy(), l1_o.numpy(), rtol=3e-6, atol=1e-4)
    (l1.sum() / img.numel()).backward()
    for k, (p, op) in enumerate(zip(ps, ops)):
        if op in (7, -1):
            continue
        assert rel_err(p.grad.cpu(), gp_o[k]) <= TOL_GRAD, 'param grad of op %d at position %d' % (op, k)
    assert rel_err(x.grad.cpu(), gi_o) <= TOL_GRAD


@pytest.mark.parametrize('mask_ch', [1, 3])
@pytest.mark.parametrize('ops', [[0, 1, 6], [6, 5], [2, 3], [7, 1]], ids=lambda c: '-'.join(map(str, c)))
def test_chain_with_mask(TF, ops, mask_ch):
    B, H, W = 2, 36, 52
    g = torch.Generator().manual_seed(99 + mask_ch + sum(ops))
    img = torch.rand(B, 3, H, W, generator=g)
    mask = (torch.rand(B, mask_ch, H, W, generator=g) > 0.5).float() if mask_ch == 1 else torch.rand(B, 3, H, W, generator=g)
    params = [sample_params(op, B, g) for op in ops]
    target = torch.rand(B, 3, H, W, generator=g)
    out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, ops, params, target, mask)
    x = img.cuda().requires_grad_()
    ps = [p.cuda().requires_grad_() for p in params]
    out = TF.chain(x, ops, ps, mask.cuda())
    assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
    l1 = TF.chain_l1(x, ops, ps, target.cuda(), mask.cuda())
    (l1.sum() / img.numel()).backward()
    for k, (p, op) in enumerate(zip(ps, ops)):
        if op != 7:
            assert rel_err(p.grad.cpu(), gp_o[k]) <= TOL_GRAD
    assert rel_err(x.grad.cpu(), gi_o) <= TOL_GRAD


def test_gray_and_saturated_pixels_gradients(TF):
    """Exact gray / black / white pixels (real photographs have blown highlights): the HSV operators
    route the whole gradient through channel 0 there, like autograd on the reference graph."""
    vals = torch.tensor([[0., 0, 0], [1, 1, 1], [.5, .5, .5], [.25, .25, .25], [.7, .2, .4]])
    img = vals.t().reshape(1, 3, 1, 5).repeat(1, 1, 4, 1).contiguous()
    g = torch.Generator().manual_seed(5)
    wgt = torch.randn(1, 3, 4, 5, generator=g)
    for op in (0, 1, 2, 3, 5):
        for pv in (0.3, -0.4):
            p = torch.full((1, O.num_params(op)), pv if op not in (3, 5) else 1.0 + pv)
            out_o, _, gp_o, gi_o = oracle_chain_with_grads(img, [op], [p], None, None, wgt)
            x = img.cuda().requires_grad_()
            pc = p.cuda().requires_grad_()
            out = TF.chain(x, [op], [pc])
            (out * wgt.cuda()).sum().backward()
            assert max_abs(out.detach().cpu(), out_o) <= TOL_PIX
            assert max_abs(x.grad.cpu(), gi_o) <= 1e-5, 'op %d p %.1f' % (op, pv)
            assert rel_err(pc.grad.cpu(), gp_o[0]) <= TOL_GRAD


def test_l1_sum_and_get_dist(TF):
    from t2onet_b200 import planner
    g = torch.Generator().manual_seed(3)
    for shape in [(1, 3, 128, 128), (4, 3, 33, 7), (2, 3, 600, 901)]:
        a, b = torch.rand(*shape, generator=g), torch.rand(*shape, generator=g)
        ref = (a - b).abs().flatten(1).sum(1)
        got = TF.l1_sum(a.cuda(), b.cuda()).cpu()
        assert np.allclose(got.numpy(), ref.numpy(), rtol=3e-6)
        d = planner.get_dist(a.cuda(), b.cuda(), 'L1')
        assert d.dim() == 0 and abs(d.item() - O.l1_dist(a, b).item()) <= TOL_PIX


def test_determinism_bitwise(TF):
    g = torch.Generator().manual_seed(11)
    img = torch.rand(4, 3, 96, 160, generator=g).cuda()
    tgt = torch.rand(4, 3, 96, 160, generator=g).cuda()
    ops = [0, 1, 2, 3, 5, 6]
    params = [sample_params(op, 4, g).cuda() for op in ops]
    r1 = TF.chain_forward_backward(img, ops, params, tgt, want_grad_img=True)
    r2 = TF.chain_forward_backward(img, ops, params, tgt, want_grad_img=True)
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1]) and torch.equal(r1[3], r2[3])
    for a, b in zip(r1[2], r2[2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize('ops', [[0, 1, 2, 3, 5, 6], [0, 1, 2, 3, 5]], ids=['c6', 'p5'])
@pytest.mark.parametrize('shape', [(2, 40, 64), (1, 96, 260), (3, 128, 128)])
def test_specialized_chain_kernels_match_generic_and_oracle(TF, ops, shape, monkeypatch):
    """The chain-specialised instantiations (t2o_step.cu: SP_C6 / SP_P5) against the run-time dispatched kernels
    (T2O_NO_SPECIALIZED=1, read at every launch) and against the oracle."""
    B, H, W = shape
    g = torch.Generator().manual_seed(79 + H + len(ops))
    img = torch.rand(B, 3, H, W, generator=g)
    params = [sample_params(op, B, g) for op in ops]
    with torch.no_grad():
        target = O.chain(img, ops, [sample_params(op, B, g) for op in ops])
    out_o, l1_o, gp_o, gi_o = oracle_chain_with_grads(img, ops, params, target)
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('T2O_NO_SPECIALIZED', mode)
        res[mode] = TF.chain_forward_backward(img.cuda(), ops, [p.cuda() for p in params], target.cuda(), want_grad_img=True)
        torch.cuda.synchronize()
    for mode, (out, l1, grads, gimg) in res.items():
        assert max_abs(out.cpu(), out_o) <= TOL_PIX
        assert np.allclose(l1.cpu().numpy(), l1_o.numpy(), rtol=3e-6, atol=1e-4)
        # a pixel whose forward value sits within an ulp of a kink (clamp edge, curve knot, out == target of the L1)
        # moves a mean-L1 parameter gradient by ~1/numel * |dy/dp|; the kink pixels are counted from the image
        # gradients (<= 2 allowed) and only they buy absolute slack -- a kink-free run is held to TOL_GRAD as is
        slack = kink_slack(gimg.cpu(), gi_o, img.numel())
        for k in range(len(ops)):
            assert rel_err(grads[k].cpu(), gp_o[k], atol=slack) <= TOL_GRAD, (mode, k)
        assert rel_err_kinks(gimg.cpu(), gi_o) <= TOL_GRAD, mode
    (o0, l0, g0, i0), (o1, l1_, g1, i1) = res['0'], res['1']
    assert max_abs(o0.cpu(), o1.cpu()) <= 1e-6 and rel_err_kinks(i0.cpu(), i1.cpu()) <= 1e-5
    for a, b in zip(g0, g1):
        assert rel_err(a.cpu(), b.cpu()) <= 1e-5


def test_score_candidates_oracle(TF):
    g = torch.Generator().manual_seed(21)
    for (S, T, H, W) in [(3, 1, 64, 64), (2, 2, 40, 52), (1, 1, 33, 7), (2, 1, 128, 128)]:
        states = torch.rand(S, 3, H, W, generator=g)
        targets = torch.rand(T, 3, H, W, generator=g)
        cs, co, cp = [], [], []
        for s in range(S):
            for op in [0, 1, 2, 3, 5, 6, 8, 9, 7, 10, 11, 12]:
                for rep in range(2 if s != 1 else 1):
                    cs.append(s); co.append(op)
                    row = torch.zeros(24); p = sample_params(op, 1, g)[0]; row[:p.numel()] = p
                    cp.append(row)
        cp = torch.stack(cp)
        got = TF.score_candidates(states.cuda(), targets.cuda(), cs, co, cp).cpu().numpy()
        for c in range(len(cs)):
            n = O.num_params(co[c])
            ref = (O.execute(co[c], states[cs[c]:cs[c] + 1], cp[c:c + 1, :n]) - targets[cs[c] % T]).abs().sum().item()
            assert abs(got[c] - ref) <= max(3e-6 * ref, 1e-4), (S, T, H, W, c, co[c], got[c], ref)
    assert TF.score_candidates(states.cuda(), targets.cuda(), [], [], torch.zeros(0, 24)).numel() == 0


def test_identity_chain_and_large_image_properties(TF):
    """Size-independent properties at a high-resolution shape (2 x 3 x 2048 x 3072): neutral parameters
    leave the image unchanged, the fused L1 of the neutral chain is ~0, the L1 is symmetric, and
    forward-only and fused forward+backward agree bit for bit."""
    B, H, W = 2, 2048, 3072
    g = torch.Generator(device='cuda').manual_seed(4)
    img = torch.rand(B, 3, H, W, generator=g, device='cuda')
    ops = [0, 1, 2, 3, 5, 6]
    neutral = [torch.zeros(B, 1), torch.zeros(B, 1), torch.zeros(B, 1), torch.ones(B, 24), torch.ones(B, 8), torch.zeros(B, 1)]
    neutral = [p.cuda() for p in neutral]
    out = TF.chain(img, ops, neutral)
    assert (out - img).abs().max().item() <= 5e-6
    l1 = TF.chain_l1(img, ops, neutral, img)
    assert (l1 / (3 * H * W)).max().item() <= 1e-6
    params = [sample_params(op, B, torch.Generator().manual_seed(8)).cuda() for op in ops]
    tgt = torch.rand(B, 3, H, W, generator=g, device='cuda')
    out1 = TF.chain(img, ops, params)
    l1a = TF.chain_l1(img, ops, params, tgt)
    out2, l1b, grads, _ = TF.chain_forward_backward(img, ops, params, tgt)
    assert (out1 - out2).abs().max().item() <= 1e-6
    assert np.allclose(l1a.cpu().numpy(), l1b.cpu().numpy(), rtol=2e-6)
    ref = (out1 - tgt).abs().flatten(1).sum(1, dtype=torch.float64)
    assert np.allclose(l1a.double().cpu().numpy(), ref.cpu().numpy(), rtol=2e-6)
    assert np.allclose(TF.l1_sum(out1, tgt).cpu().numpy(), TF.l1_sum(tgt, out1).cpu().numpy(), rtol=0, atol=0)
    for gk in grads:
        assert torch.isfinite(gk).all()
    # linearity of the parameter gradient in the loss scale
    _, _, g2, _ = TF.chain_forward_backward(img, ops, params, tgt, want_out=False,
                                            loss_scale=torch.full((B,), 2.0 / img.numel(), device='cuda'))
    for a, b in zip(grads, g2):
        assert np.allclose((2 * a).cpu().numpy(), b.cpu().numpy(), rtol=1e-6, atol=1e-12)


def test_full_resolution_chain_against_the_oracle_on_the_gpu(TF):
    """BASELINE config 4's image size (3 x 2048 x 3072, one image): the fused forward + L1 + backward launch against the
    oracle's torch ops run on the same GPU in fp32 (autograd through all six operators, ~20 GB of saved planes) --
    pixels max-abs 1e-5, L1 1e-5, parameter gradients relative 1e-4, at the size the benchmark runs."""
    B, H, W = 1, 2048, 3072
    g = torch.Generator(device='cuda').manual_seed(11)
    img = torch.rand(B, 3, H, W, generator=g, device='cuda')
    ops = [0, 1, 2, 3, 5, 6]
    cg = torch.Generator().manual_seed(12)
    params = [sample_params(op, B, cg).cuda() for op in ops]
    with torch.no_grad():
        tgt = O.chain(img, ops, [sample_params(op, B, cg).cuda() for op in ops])
    ps = [p.clone().requires_grad_() for p in params]
    out_o = O.chain(img, ops, ps)
    loss_o = (out_o - tgt).abs().mean()
    loss_o.backward()
    out, l1, grads, _ = TF.chain_forward_backward(img, ops, params, tgt)
    assert (out - out_o.detach()).abs().max().item() <= TOL_PIX
    assert abs(l1.sum().item() / img.numel() - loss_o.item()) <= 1e-5
    for k, (gk, p) in enumerate(zip(grads, ps)):
        assert rel_err(gk.cpu(), p.grad.cpu()) <= TOL_GRAD, (ops[k], gk, p.grad)
    del out_o, loss_o, ps
    torch.cuda.empty_cache()


def test_errors_are_loud(TF):
    from t2onet_b200 import T2OError
    img = torch.rand(1, 3, 8, 8).cuda()
    with pytest.raises(T2OError):
        TF.chain(torch.rand(1, 3, 8, 8), [0], [torch.zeros(1, 1)])            # CPU tensor: no fallback
    with pytest.raises(T2OError):
        TF.chain(img, [4], [torch.zeros(1, 1).cuda()])                        # inpaint unsupported
    with pytest.raises(T2OError):
        TF.chain(img, [3], [torch.zeros(1, 3).cuda()])                        # too few parameters
    with pytest.raises(T2OError):
        TF.chain(img.double(), [0], [torch.zeros(1, 1).cuda()])


def test_random_chains_shapes_and_masks_sweep(TF):
    """A seeded sweep over random shapes (ragged widths -> 4-, 2- and 1-pixel groups, images shorter than a pipeline
    band, single rows / columns), random operator chains over all twelve operators (repeats and identities included:
    the binding splits them into launches) and random masks: pixels, L1 and gradients against the oracle."""
    import random
    rnd = random.Random(2024)
    pool = [0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, -1]
    for trial in range(40):
        B = rnd.choice([1, 2, 3])
        H = rnd.choice([1, 2, 5, 9, 17, 33, 64, 100])
        W = rnd.choice([1, 3, 4, 6, 10, 30, 64, 65, 128, 130, 200])
        n = rnd.randint(1, 7)
        ops = [rnd.choice(pool) for _ in range(n)]
        if all(o < 0 for o in ops):
            ops[0] = 1
        mask_ch = rnd.choice([0, 0, 1, 3])
        g = torch.Generator().manual_seed(1000 + trial)
        img = torch.rand(B, 3, H, W, generator=g)
        params = [sample_params(op, B, g) if op >= 0 else torch.zeros(B, 1) for op in ops]
        mask = None if mask_ch == 0 else (torch.rand(B, mask_ch, H, W, generator=g) > 0.3).float()
        target = torch.rand(B, 3, H, W, generator=g)
        xo = img.clone().requires_grad_()
        po = [p.clone().requires_grad_() for p in params]
        out_o = O.chain(xo, ops, po, mask)
        loss_o = (out_o - target).abs().mean()
        loss_o.backward()
        x = img.cuda().requires_grad_()
        ps = [p.cuda().requires_grad_() for p in params]
        out = TF.chain(x, ops, ps, None if mask is None else mask.cuda())
        tag = (trial, B, H, W, ops, mask_ch)
        assert max_abs(out.detach().cpu(), out_o.detach()) <= TOL_PIX, tag
        loss = (out - target.cuda()).abs().mean()
        loss.backward()
        assert abs(loss.item() - loss_o.item()) <= 1e-5, tag
        # kink pixels (a forward value within an ulp of a clamp edge / curve knot / the target) may take the other
        # one-sided derivative: counted exactly from the image gradient, and the parameter gradients get their slack
        slack = kink_slack(x.grad.cpu(), xo.grad, img.numel(), max_kinks=3)
        for k, (p, q) in enumerate(zip(ps, po)):
            if ops[k] < 0 or ops[k] == 7 or q.grad is None:
                continue
            d = (p.grad.cpu() - q.grad).abs().max().item()
            assert d <= max(TOL_GRAD * q.grad.abs().max().item(), slack), (tag, k, p.grad, q.grad)
