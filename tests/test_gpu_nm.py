"""Device-resident Nelder-Mead (t2o_nm_start / t2o_nm_advance, csrc/t2o_nm.cu) against the host coroutine restating
scipy's _minimize_neldermead (t2onet_b200/nelder_mead.py, itself checked against scipy in test_host_logic.py): both are
driven by the same scorer (t2o_score_candidates), so equal function values must give the same vertex sequence --
final vertices, function values, iteration and evaluation counts are compared EXACTLY.  (The coroutine runs with
stable=True: exact ties between fp32 function values are common, and scipy's tie order is numpy's unstable-quicksort
implementation detail, which the device's stable rank sort does not imitate.)"""
import os

import numpy as np
import pytest
import torch

from oracle import ops as O
from parity_util import sample_params

pytestmark = pytest.mark.gpu
GLOBAL_OPS = [0, 1, 2, 3, 5, 6]


@pytest.fixture(scope='module')
def T():
    import t2onet_b200 as T
    return T


def _pairs(S, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(S, 3, H, W, generator=g)
    tgt = img.clone()
    for op in (0, 5, 6):                                   # a planted edit so that the fits have something to find
        tgt = O.execute(op, tgt, sample_params(op, S, g))
    return img.cuda(), tgt.cuda()


def test_device_nm_equals_host_coroutine(T):
    from t2onet_b200 import planner
    ex = T.Executor(T.default_options()).cuda()
    states, targets = _pairs(3, 32, 48, 77)
    problems = [(0, 0), (0, 1), (0, 2), (0, 5), (0, 6), (1, 3), (1, 5), (1, 6), (2, 0), (2, 2)]
    host = planner.fit_params_nelder_mead_host(states, targets, problems, ex, state_target=[0, 1, 2], stable=True)
    dev = planner.fit_params_nelder_mead(states, targets, problems, ex, state_target=[0, 1, 2])
    for (s, op), h, d in zip(problems, host, dev):
        assert d.nfev == h.nfev and d.nit == h.nit and d.status == h.status, (s, op, d.nfev, h.nfev, d.nit, h.nit)
        assert np.array_equal(np.asarray(d.x), np.asarray(h.x)), (s, op, d.x, h.x)
        assert d.fun == h.fun, (s, op, d.fun, h.fun)


def test_device_nm_scalar_fit_finds_planted_parameter(T):
    from t2onet_b200 import planner
    ex = T.Executor(T.default_options()).cuda()
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 40, 64, generator=g)
    p = torch.tensor([[0.23]])
    tgt = O.execute(0, img, p)
    param, ok = planner.get_param(img.cuda(), tgt.cuda(), None, 0, ex, None, 'L1', 'Nelder-Mead')
    assert ok and param.dtype == torch.float64 and tuple(param.shape) == (1, 1)
    assert abs(param.item() - 0.23) < 2e-3


def test_beam_search_batch_equals_single_pairs(T):
    """beam_search_batch over M pairs == beam_search on each pair alone (the lock-step only shares launches)."""
    from t2onet_b200 import planner
    ex = T.Executor(T.default_options()).cuda()
    img, tgt = _pairs(3, 32, 32, 9)
    batch = planner.beam_search_batch(img, tgt, ex, 2, [0, 1, 6], O.ACTION_NAMES, 2, 1e-3)
    for m in range(3):
        actions, Is = planner.beam_search(img[m:m + 1], tgt[m:m + 1], None, ex, None, 2, [0, 1, 6], O.ACTION_NAMES, 2, 1e-3,
                                          'L1', 'Nelder-Mead')
        b_actions, b_Is = batch[m]
        assert [[a[0] for a in seq] for seq in actions] == [[a[0] for a in seq] for seq in b_actions]
        for seq, bseq in zip(actions, b_actions):
            for a, b in zip(seq, bseq):
                assert a[1] == b[1] and a[2] == b[2]
        for seq, bseq in zip(Is, b_Is):
            for x, y in zip(seq, bseq):
                assert not x.is_cuda and torch.equal(x, y)


@pytest.mark.parametrize('resident', ['1', '0'])
def test_beam_search_pipelined_equals_sequential_batches(T, resident, monkeypatch):
    """planner.beam_search_pipelined (batches in flight on two threads / CUDA streams) == one beam_search_batch per batch:
    same sequences, parameters, distances, images and evaluation count -- with the resident Nelder-Mead launches and with the
    round-by-round fallback (whose CUDA-graph captures must not collide between the threads)."""
    from t2onet_b200 import planner
    monkeypatch.setenv('T2O_NM_RESIDENT', resident)
    ex = T.Executor(T.default_options()).cuda()
    batches = [_pairs(3 + (k % 2), 32, 32, 40 + k) for k in range(5)]
    c_seq, c_pipe = [0], [0]
    seq = [planner.beam_search_batch(a, b, ex, 2, [0, 1, 3, 6], O.ACTION_NAMES, 2, 1e-3, counter=c_seq) for a, b in batches]
    pipe = planner.beam_search_pipelined(iter(batches), ex, 2, [0, 1, 3, 6], O.ACTION_NAMES, 2, 1e-3, workers=2, counter=c_pipe)
    top = planner.beam_search_pipelined(iter(batches), ex, 2, [0, 1, 3, 6], O.ACTION_NAMES, 2, 1e-3, workers=2, images='top')
    for rs, rt in zip(seq, top):
        for (acts, Is), (tacts, tIs) in zip(rs, rt):
            assert acts == tacts and len(tIs) == len(Is) and all(len(x) == 0 for x in tIs[1:])
            assert len(tIs[0]) == len(Is[0]) and all(torch.equal(x, y) for x, y in zip(Is[0], tIs[0]))
    assert c_seq[0] == c_pipe[0] and len(seq) == len(pipe)
    for rs, rp in zip(seq, pipe):
        assert len(rs) == len(rp)
        for (acts, Is), (pacts, pIs) in zip(rs, rp):
            assert acts == pacts
            for xs, ys in zip(Is, pIs):
                assert len(xs) == len(ys) and all(torch.equal(x, y) for x, y in zip(xs, ys))


def test_entry_points_share_one_workspace_across_batch_sizes(T):
    """include/t2o.h: one workspace serves all entry points in turn, whatever the batch sizes -- the arrival counters
    live in a fixed region that no call's partial sums can reach (a scorer launch with few states followed by a
    chain launch over many rows, and the other way round, used to overlap them)."""
    from t2onet_b200 import _lib, functional as TF
    states, targets = _pairs(2, 32, 48, 3)
    prm = torch.rand(20, 24) + 0.5
    sc0 = TF.score_candidates(states, targets, [0] * 10 + [1] * 10, [5] * 20, prm).clone()
    big, big_t = _pairs(40, 32, 48, 4)
    for _ in range(2):
        out, l1 = TF._rows_forward_raw(*TF._prep_row_ops([[0]] * 40, 40, big.device), big, None, 0,
                                       torch.full((40, 24), 0.1, device=big.device), big_t, True, True, 8)
        ref = (out - big_t).abs().flatten(1).sum(1)
        assert torch.allclose(l1, ref, rtol=1e-5)
        many, many_t = _pairs(30, 32, 48, 6)
        sc = TF.score_candidates(many, many_t, list(range(30)), [0] * 30, torch.full((30, 24), 0.1))
        ref = (TF.execute_rows(many, [[0]] * 30, torch.full((30, 24), 0.1, device=many.device)) - many_t).abs().flatten(1).sum(1)
        assert torch.allclose(sc, ref, rtol=1e-5)
        assert torch.equal(TF.score_candidates(states, targets, [0] * 10 + [1] * 10, [5] * 20, prm), sc0)
    torch.cuda.synchronize()
    ws = _lib.workspace(states.device, 1)
    assert not bool(ws[:65536 * 4].any())


def test_topk_min_is_the_stable_argsort_prefix(T):
    """t2o_topk_min == np.argsort(kind='stable')[:k] per segment (utils/beam_search.py:252-256), with ties, NaN, short and
    empty segments."""
    import t2onet_b200.functional as TF
    g = torch.Generator().manual_seed(9)
    sizes = [0, 1, 3, 8, 17, 64, 200, 1000]
    vals = []
    for n in sizes:
        v = torch.rand(n, generator=g)
        if n >= 8:
            v[::3] = v[0]                                   # exact ties
            v[5] = float('nan')
        vals.append(v)
    values = torch.cat(vals)
    seg = np.concatenate([[0], np.cumsum(sizes)])
    for k in (1, 8, 32):
        idx, val = TF.topk_min(values.cuda(), seg, k)
        idx, val = idx.cpu().numpy(), val.cpu().numpy()
        for s, n in enumerate(sizes):
            v = values[seg[s]:seg[s + 1]].numpy()
            order = np.argsort(np.where(np.isnan(v), np.inf, v), kind='stable')[:k]
            assert list(idx[s, :len(order)]) == [int(o) + int(seg[s]) for o in order], (k, s)
            assert all(i == -1 for i in idx[s, len(order):]) and np.all(np.isinf(val[s, len(order):]))
            got = val[s, :len(order)]
            assert np.array_equal(np.isnan(got), np.isnan(v[order])) and np.array_equal(got[~np.isnan(got)], v[order][~np.isnan(v[order])])


@pytest.mark.parametrize('S,H,W,masked', [(5, 128, 128, False), (80, 128, 128, False), (6, 64, 96, True), (4, 32, 32, False),
                                          (3, 40, 52, False), (2, 256, 256, False), (2, 264, 320, True), (4, 64, 128, 3)])
def test_resident_nelder_mead_equals_the_rounds_exactly(T, S, H, W, masked, monkeypatch):
    """t2o_nm_run_resident (a cluster of CTAs keeps each state in shared memory for the life of its fits) against rounds of
    t2o_score_candidates + t2o_nm_advance: fitted parameters, function values, iteration and evaluation counts identical,
    with one to six fits per state, several tiles per image, ragged tiles and masks (256 x 256 has 16 tiles: a cluster of 16;
    264 x 320 has 27: the resident path declines and the rounds run)."""
    from t2onet_b200 import planner
    ex = T.Executor(T.default_options()).cuda()
    states, targets = _pairs(S, H, W, 31 + S)
    ops = [0, 1, 2, 6] if S > 10 else GLOBAL_OPS
    problems = [(s, op) for s in range(S) for op in ops[:1 + (s * 5) % len(ops)]]
    kw = {}
    if masked:
        g = torch.Generator().manual_seed(3)
        kw = dict(masks=(torch.rand(2, 3 if masked == 3 else 1, H, W, generator=g) > 0.4).float().cuda(),
                  prob_mask=[(i % 3) - 1 for i in range(len(problems))])
    res = []
    for env in ('0', '1'):
        monkeypatch.setenv('T2O_NM_RESIDENT', env)
        res.append(planner.fit_params_nelder_mead(states, targets, problems, ex, state_target=list(range(S)), **kw))
    for (s, op), r0, r1 in zip(problems, *res):
        assert r0.nfev == r1.nfev and r0.nit == r1.nit and r0.status == r1.status, (s, op, r0.nfev, r1.nfev)
        assert np.array_equal(np.asarray(r0.x), np.asarray(r1.x)) and r0.fun == r1.fun, (s, op)
    assert max(r.nfev for r in res[0]) >= (100 if S <= 10 else 20)


def test_resident_nelder_mead_can_stop_and_go_on_in_rounds(T):
    """max_rounds stops the resident launch early: the fits' state is written back, a second resident launch or the rounds carry
    on from there, and the results equal an uninterrupted run exactly."""
    from t2onet_b200 import planner, functional as TF
    ex = T.Executor(T.default_options()).cuda()
    S = 6
    states, targets = _pairs(S, 128, 128, 77)
    probs = [(s, op) for s in range(S) for op in GLOBAL_OPS]

    def make():
        return TF.DeviceNelderMead(states, targets, [p[0] for p in probs], [p[1] for p in probs],
                                   [planner._param0(p[1], ex) for p in probs], state_target=list(range(S)))
    ref = make()
    assert ref.run_resident()
    r0 = ref.result()
    assert bool(r0['done'].all())
    a = make()
    assert a.run_resident(37) and a.active() > 0
    assert a.run_resident(21)
    os.environ['T2O_NM_RESIDENT'] = '0'
    try:
        r1 = a.run()
    finally:
        os.environ.pop('T2O_NM_RESIDENT', None)
    for k in ('x', 'fun', 'nit', 'nfev', 'status'):
        assert torch.equal(r0[k], r1[k]), k


def test_resident_nelder_mead_declines_what_it_cannot_hold(T):
    """More than eight fits on one state (or a state without fits next to it): run_resident() answers False / skips, run() falls
    back to the rounds, and the results equal the host coroutine-checked round path."""
    from t2onet_b200 import planner, functional as TF
    ex = T.Executor(T.default_options()).cuda()
    states, targets = _pairs(3, 32, 64, 5)
    nine = [0, 1, 2, 6, 0, 1, 2, 6, 0]
    probs = [(0, op) for op in nine] + [(2, 0)]                 # state 1 has no fit at all
    nm = TF.DeviceNelderMead(states, targets, [p[0] for p in probs], [p[1] for p in probs],
                             [planner._param0(p[1], ex) for p in probs], state_target=[0, 1, 2])
    assert nm.run_resident() is False
    r = nm.run()
    assert bool(r['done'].all())
    # the duplicated problems are the same fit
    assert torch.equal(r['x'][0], r['x'][4]) and torch.equal(r['x'][0], r['x'][8]) and r['nfev'][0] == r['nfev'][4]
    few = [(0, 0), (2, 6)]
    nm2 = TF.DeviceNelderMead(states, targets, [p[0] for p in few], [p[1] for p in few], [planner._param0(p[1], ex) for p in few],
                              state_target=[0, 1, 2])
    assert nm2.run_resident() is True
    r2 = nm2.result()
    assert bool(r2['done'].all()) and torch.equal(r2['x'][0], r['x'][0]) and r2['fun'][0] == r['fun'][0]
