"""GIER-shaped planning (BASELINE config 5): masks in the candidate scorer and the planner entry
point with the arguments the GIER driver passes (preprocess/gen_greedy_seqs_GIER.py:60-71: a list of masks, the global
all-ones one first, and the operator each local mask belongs to).  The reference's own beam_search does not accept these
arguments (its GIER driver does not run as committed), so the checker is the oracle: Operator.execute's blend
(models/operators.py:129-130) for the scores, and oracle/planner.py's beam_search with the same candidate enumeration --
scipy's Nelder-Mead around the oracle operators on the CPU -- for the search."""
import pytest
import torch

from oracle import ops as O
from oracle import planner as OP
from parity_util import sample_params

pytestmark = pytest.mark.gpu
NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']


@pytest.fixture(scope='module')
def T():
    import t2onet_b200 as T
    return T


def _masks(n, ch, H, W, g):
    """soft-edged rectangles (values in [0, 1], exact 0 / 1 regions included)"""
    m = torch.zeros(n, ch, H, W)
    for i in range(n):
        y0, x0 = int(torch.randint(0, H // 2, (1,), generator=g)), int(torch.randint(0, W // 2, (1,), generator=g))
        m[i, :, y0:y0 + H // 2, x0:x0 + W // 2] = 1.0
        m[i, :, y0:y0 + 2, x0:x0 + W // 2] = 0.5
        if ch == 3:
            m[i, 1] *= 0.75
    return m


@pytest.mark.parametrize('H,W,ch', [(128, 128, 1), (40, 52, 3), (33, 30, 1), (256, 256, 3)])
def test_masked_scorer_vs_oracle(T, H, W, ch):
    import t2onet_b200.functional as TF
    g = torch.Generator().manual_seed(H * 7 + W + ch)
    S = 6
    states = torch.rand(S, 3, H, W, generator=g)
    targets = torch.rand(2, 3, H, W, generator=g)
    masks = _masks(3, ch, H, W, g)
    ops_all = [0, 1, 2, 3, 5, 6, 8, 9, 10, 11, 12]
    cand_state, cand_op, cand_mask, prm, ref = [], [], [], [], []
    for s in range(S):
        n = 3 + s * 3                                        # 3 .. 18 candidates: both of the scorer's candidate paths
        for c in range(n):
            op = ops_all[(s * 5 + c) % len(ops_all)]
            mk = (s + c) % 4 - 1                             # -1 (no mask), 0, 1, 2
            p = sample_params(op, 1, g)
            row = torch.zeros(24)
            row[:p.shape[1]] = p[0]
            cand_state.append(s); cand_op.append(op); cand_mask.append(mk); prm.append(row)
            out = O.execute(op, states[s:s + 1], p, None if mk < 0 else masks[mk:mk + 1].expand(1, 3, H, W))
            ref.append(float((out - targets[s % 2:s % 2 + 1]).abs().double().sum()))
    got = TF.score_candidates(states.cuda(), targets.cuda(), cand_state, cand_op, torch.stack(prm), masks=masks.cuda(),
                              cand_mask=cand_mask).cpu().double()
    ref = torch.tensor(ref, dtype=torch.float64)
    assert float(((got - ref).abs() / ref).max()) <= 2e-5
    # a launch that carries masks scores its unmasked candidates exactly like a launch without masks
    plain = TF.score_candidates(states.cuda(), targets.cuda(), cand_state, cand_op, torch.stack(prm)).cpu()
    nomask = TF.score_candidates(states.cuda(), targets.cuda(), cand_state, cand_op, torch.stack(prm), masks=masks.cuda(),
                                 cand_mask=[-1] * len(cand_op)).cpu()
    assert torch.equal(plain, nomask)
    sel = torch.tensor([m < 0 for m in cand_mask])
    assert torch.equal(plain[sel], got.float()[sel])


def _gier_pair(H, W, seed):
    """Planted GIER-style edit: a local brightness change inside mask 1, a global contrast change, a local saturation change
    inside mask 2."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(1, 3, 6, 6, generator=g)
    I0 = (torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False) * 0.6 + 0.2 +
          (torch.rand(1, 3, H, W, generator=g) - 0.5) * 0.1).clamp(0.02, 0.98)
    m1 = torch.zeros(1, 1, H, W); m1[..., H // 8:H // 2, W // 8:W // 2] = 1.0
    m2 = torch.zeros(1, 1, H, W); m2[..., H // 2:, W // 3:] = 1.0
    masks = [torch.ones(1, 3, H, W), m1.expand(1, 3, H, W).clone(), m2.expand(1, 3, H, W).clone()]
    mask_op_idx = [-1, 0, 2]
    x = O.execute(0, I0, torch.tensor([[0.25]]), masks[1])
    x = O.execute(1, x, torch.tensor([[0.3]]), None)
    Igt = O.execute(2, x, torch.tensor([[0.35]]), masks[2])
    return I0, Igt, masks, mask_op_idx


@pytest.mark.parametrize('H,W', [(32, 32), (256, 256)])
def test_beam_search_gier_vs_oracle_planner(T, H, W):
    I0, Igt, masks, mask_op_idx = _gier_pair(H, W, 77 + H)
    ops, beam, max_step, err = [0, 1, 2, 6], 3, 3, 1e-3
    ex = T.Executor(T.default_options()).cuda()
    trace = []
    res = T.planner.beam_search_batch(I0.cuda(), Igt.cuda(), ex, beam, ops, NAMES, max_step, err, trace=trace,
                                      masks=[[m.cuda() for m in masks]], mask_op_idx=[mask_op_idx])
    actions, Is = res[0]
    otrace = []
    o_actions, _ = OP.beam_search(I0, Igt, None, O.OracleExecutor(), None, beam, ops, NAMES, max_step, err, 'L1', 'Nelder-Mead',
                                  trace=otrace, mask=masks, mask_op_idx=mask_op_idx)
    # same candidates at every step, the same (operator, mask) sequences kept, distances within the fit tolerance
    assert len(trace[0]['steps']) == len(otrace)
    for st, ost in zip(trace[0]['steps'], otrace):
        a = [(c['parent'], c['op'], c['mask']) for c in st['candidates']]
        b = [(c['parent'], c['op'], c['mask']) for c in ost['candidates']]
        assert a == b
        for c, oc in zip(st['candidates'], ost['candidates']):
            assert abs(c['dist'] - oc['dist']) <= 2e-4, (c, oc)
    assert [[(a[0], a[3]) for a in seq] for seq in actions] == [[(a[0], a[3]) for a in seq] for seq in o_actions]
    # the planted edit is found: brightness inside mask 1, global contrast, saturation inside mask 2
    top = sorted((a[0], a[3]) for a in actions[0])
    assert top == [('brightness', 1), ('contrast', 0), ('saturation', 2)], actions[0]
    assert actions[0][-1][2] < err
    # the public GIER-style entry point returns the same search
    a2, _ = T.planner.beam_search_gier(I0.cuda(), Igt.cuda(), None, [m.cuda() for m in masks], mask_op_idx, ex, beam, ops, NAMES,
                                       max_step, err, 'L1', 'Nelder-Mead')
    assert [[(a[0], a[3]) for a in seq] for seq in a2] == [[(a[0], a[3]) for a in seq] for seq in actions]
    # replaying the top sequence through the Executor with its masks reproduces the returned images
    img = I0.cuda()
    for a, I_k in zip(actions[0], Is[0]):
        mk = None if mask_op_idx[a[3]] < 0 else masks[a[3]].cuda()
        img = ex.execute(img, NAMES.index(a[0]), mk, specified_param=torch.tensor([a[1]], device='cuda'))[0]
        assert float((img.cpu() - I_k).abs().max()) <= 1e-5
