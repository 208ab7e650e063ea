"""Drop-in operation planner: utils/beam_search.py of the reference (and its fixed-order /
eps-greedy variants) with the same function names, signatures and return structure, running on the
candidate-scoring kernel.

What changes underneath: the reference fits the parameters of every
(beam state, operator) pair one after the other, each Nelder-Mead evaluation being an
`executor.execute` + `get_dist` + `.item()` round trip (utils/beam_search.py:77-87).  Here all fits
of a beam step -- of every image pair in flight (`beam_search_batch`) -- advance in lock-step as
device-resident Nelder-Mead state machines (t2o_nm_start / t2o_nm_advance, csrc/t2o_nm.cu); a round
scores the pending vertex of every fit with ONE `t2o_score_candidates` launch over the TMA-staged state
tiles, and vertices / scores never leave device memory (rounds replay as a CUDA graph).  The fitted
candidates of a step are applied and scored by one per-row launch (t2o_rows_forward).  The selection
logic (utils/beam_search.py:239-259) is unchanged.  t2onet_b200/nelder_mead.py is the same
Nelder-Mead as a host coroutine (`T2O_HOST_NM=1` or `fit_params_nelder_mead_host`), kept as the
cross-check of the device version.

Only the 'L1' distance is implemented: the discriminator branches of the reference call undefined
names (utils/beam_search.py:42,55) and are dead code.
"""
import os
import threading
import random

import numpy as np
import torch

from . import functional as TF
from .nelder_mead import nelder_mead, run_lockstep

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
NM_INIT_ZERO, NM_INIT_ONE = (0, 1, 2, 6), (3, 5)
# utils/beam_search_eps_greedy.py:24 seeds the global `random` module at import (random.seed(0)) and draws from it
# (:299-300).  The drop-in keeps its own generator in the same state instead of re-seeding the caller's global one:
# the draws are identical as long as nothing else consumes the reference's global stream.
_eps_rng = random.Random(0)


def eps_greedy_seed(seed=0):
    """Re-seed the eps-greedy planner's generator (what re-importing utils/beam_search_eps_greedy.py does)."""
    _eps_rng.seed(seed)


def get_dist(x1, x2, dist_type='L1'):
    """utils/beam_search.py:170-180: (x1 - x2).norm(1) / x1.numel() as a 0-dim tensor."""
    if dist_type != 'L1':
        assert False, '{} is invalid distance'.format(dist_type)
    if x1.requires_grad or x2.requires_grad:
        return _L1Dist.apply(x1, x2)
    return TF.l1_sum(x1, x2).sum() / x1.numel()


class _L1Dist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2):
        ctx.save_for_backward(x1, x2)
        return TF.l1_sum(x1, x2).sum() / x1.numel()

    @staticmethod
    def backward(ctx, g):
        x1, x2 = ctx.saved_tensors
        s = torch.sign(x1 - x2) * (g / x1.numel())
        return s, -s


def execute(I, operation, param, executor):
    """utils/beam_search.py:165-167"""
    img, _ = executor.execute(I, operation, None, features=None, specified_param=param, has_noise=False)
    return img


def _param0(operation, executor):
    """utils/beam_search.py:149-155"""
    n = executor.get_param_num(operation)
    if operation in NM_INIT_ZERO:
        return np.zeros(n)
    if operation in NM_INIT_ONE:
        return np.ones(n)
    assert False, 'the operation is not global operation'


def fit_params_nelder_mead_host(states, targets, problems, executor, state_target=None, counter=None, stable=False):
    """Fit many (state index, operator) problems at once with the host coroutine (one launch + one sync per round).

    states (S,3,H,W) / targets (T,3,H,W) CUDA tensors; problems: list of (state_idx, operation).
    Returns a list of NMResult in problem order (x is float64, as scipy returns it)."""
    numel = float(states[0].numel())
    L = getattr(executor.opt, 'curve_steps', 8)
    gens = {i: nelder_mead(_param0(op, executor), stable=stable) for i, (s, op) in enumerate(problems)}
    order = sorted(range(len(problems)), key=lambda i: (problems[i][0], i))   # candidates sorted by state
    rank = {k: r for r, k in enumerate(order)}

    def score(keys, points):
        keys_sorted = sorted(keys, key=lambda k: rank[k])
        pts = dict(zip(keys, points))
        C = len(keys_sorted)
        prm = np.zeros((C, 24), dtype=np.float32)
        for r, k in enumerate(keys_sorted):
            p = pts[k]
            prm[r, :len(p)] = p                     # float64 -> float32, as torch.tensor([param], dtype=torch.float)
        l1 = TF.score_candidates(states, targets, [problems[k][0] for k in keys_sorted],
                                 [problems[k][1] for k in keys_sorted], torch.from_numpy(prm),
                                 state_target=state_target, curve_steps=L)
        vals = (l1 / numel).tolist()                # fp32 division, then .item() -> python float
        if counter is not None:
            counter[0] += C
        by_key = dict(zip(keys_sorted, vals))
        return [by_key[k] for k in keys]

    res = run_lockstep(gens, score)
    return [res[i] for i in range(len(problems))]


class _Fit:
    __slots__ = ('x', 'fun', 'nit', 'nfev', 'status', 'success')

    def __init__(self, x, fun, nit, nfev, status):
        self.x, self.fun, self.nit, self.nfev, self.status = x, fun, nit, nfev, status
        self.success = status == 0


def fit_params_nelder_mead(states, targets, problems, executor, state_target=None, counter=None, numel=None,
                           masks=None, prob_mask=None):
    """Fit many (state index, operator) problems at once on the device (TF.DeviceNelderMead).
    masks (n_masks, 1|3, H, W) + prob_mask (one mask index or -1 per problem): the fit edits inside that mask.

    states (S,3,H,W) / targets (T,3,H,W) CUDA tensors; problems: list of (state_idx, operation).
    Returns a list of results (x float64 (n,), fun, nit, nfev, status, success) in problem order."""
    if os.environ.get('T2O_HOST_NM') == '1' and masks is None:
        return fit_params_nelder_mead_host(states, targets, problems, executor, state_target, counter)
    if not problems:
        return []
    L = getattr(executor.opt, 'curve_steps', 8)
    sts = [p[0] for p in problems]
    if all(a <= b for a, b in zip(sts, sts[1:])):                                # (the planner lists its problems state by state)
        order = list(range(len(problems)))
    else:
        order = sorted(range(len(problems)), key=lambda i: (sts[i], i))           # the scorer wants candidates sorted by state
    x0 = {op: _param0(op, executor) for op in set(p[1] for p in problems)}
    nm = TF.DeviceNelderMead(states, targets, [problems[i][0] for i in order], [problems[i][1] for i in order],
                             [x0[problems[i][1]] for i in order], state_target=state_target,
                             curve_steps=L, numel=numel, masks=masks,
                             prob_mask=None if masks is None else [prob_mask[i] for i in order])
    r = nm.run()
    assert bool(r['done'].all()), 'Nelder-Mead fits did not finish'
    if counter is not None:
        counter[0] += int(r['nfev'].sum())
    out = [None] * len(problems)
    x, fun = r['x'].numpy(), r['fun'].tolist()
    ns, nit, nfev, status = r['n'].tolist(), r['nit'].tolist(), r['nfev'].tolist(), r['status'].tolist()
    for pos, i in enumerate(order):
        out[i] = _Fit(x[pos, :ns[pos]].copy(), fun[pos], nit[pos], nfev[pos], status[pos])
    return out


def get_param_naive(img, out, txt, mask, param0, executor, discriminator, op_ind, dist_type, optimizer):
    """utils/beam_search.py:65-91 (L1 only): Nelder-Mead from param0 -> (param (1,n) tensor, success)."""
    assert dist_type == 'L1'
    L = getattr(executor.opt, 'curve_steps', 8)
    numel = float(img.numel())
    if os.environ.get('T2O_HOST_NM') == '1':
        gen = nelder_mead(np.asarray(param0, dtype=np.float64))

        def score(keys, points):
            prm = np.zeros((1, 24), dtype=np.float32)
            prm[0, :len(points[0])] = points[0]
            l1 = TF.score_candidates(img, out, [0], [op_ind], torch.from_numpy(prm), curve_steps=L)
            return (l1 / numel).tolist()
        res = run_lockstep({0: gen}, score)[0]
        return torch.tensor(np.array([list(res.x)])).to(img.device), res.success
    nm = TF.DeviceNelderMead(img, out, [0], [op_ind], [np.asarray(param0, dtype=np.float64)], curve_steps=L, numel=numel)
    r = nm.run()
    n = int(r['n'][0])
    return r['x'][:1, :n].clone().to(img.device), int(r['status'][0]) == 0


def gd_minimize(func, param0, method='adam'):
    """utils/beam_search.py:94-128"""
    num_iters, tol = 1000, 1e-5
    param0.requires_grad_()
    success_flag = False
    if method == 'lbfgs':
        success_flag = True
        optimizer = torch.optim.LBFGS([param0], lr=1)

        def closure():
            optimizer.zero_grad()
            loss = func(param0)
            loss.backward()
            return loss
        optimizer.step(closure)
    elif method == 'adam':
        optimizer = torch.optim.Adam([param0], lr=1e-2)
        loss_prev = 10000
        for _ in range(num_iters):
            optimizer.zero_grad()
            loss = func(param0)
            cur_loss = loss.item()
            if (loss_prev - cur_loss) < tol:
                success_flag = True
                break
            loss_prev = cur_loss
            loss.backward()
            optimizer.step()
    return param0.detach(), success_flag


def get_param_gd(img, out, txt, mask, param0, executor, discriminator, op_ind, dist_type, optimizer):
    """utils/beam_search.py:131-145: the fused chain+L1 kernel forward, the recompute kernel backward."""
    assert dist_type == 'L1'
    L = getattr(executor.opt, 'curve_steps', 8)

    def func(param):
        return TF.chain_l1(img, [op_ind], [param], out, None, L).sum() / img.numel()
    return gd_minimize(func, param0, method=optimizer)


def get_param(I0, I1, txt, operation, executor, discriminator, dist_type, optimizer):
    """utils/beam_search.py:148-162"""
    param0 = torch.from_numpy(_param0(operation, executor)).float()
    if optimizer == 'Nelder-Mead':
        return get_param_naive(I0, I1, txt, None, param0.numpy(), executor, discriminator, operation, dist_type, optimizer)
    bs = I0.shape[0]
    param0 = param0.view(1, -1).repeat(bs, 1).to(I0.device)
    return get_param_gd(I0, I1, txt, None, param0, executor, discriminator, operation, dist_type, optimizer)


def _score_outputs(I_list, ops, params, I_gt_list, executor, mask_list=None, gather=None):
    """I_out and dist for every fitted candidate of a step (utils/beam_search.py:230,237): one per-row launch
    (every row its own state, operator, parameters, target and -- GIER -- mask), one device->host read of the distances."""
    if not I_list:
        return [], []
    L = getattr(executor.opt, 'curve_steps', 8)
    dev = I_list[0].device
    numel = float(I_gt_list[0].numel())
    dev = torch.device(dev)
    outs, vals = [], []
    CH = 1024                                                       # rows per launch (bounds the temporaries)
    for c0 in range(0, len(I_list), CH):
        Is, Gs = I_list[c0:c0 + CH], I_gt_list[c0:c0 + CH]
        if gather is None:
            img = torch.cat(Is, 0).contiguous()
            tgt = torch.cat(Gs, 0).contiguous()
        else:
            # the rows are states[s_idx] / targets[pair]: two gathers instead of concatenating thousands of slices
            st_all, s_idx, gt_all, p_idx = gather
            img = st_all.index_select(0, torch.as_tensor(s_idx[c0:c0 + CH], dtype=torch.int64).to(dev))
            tgt = gt_all.index_select(0, torch.as_tensor(p_idx[c0:c0 + CH], dtype=torch.int64).to(dev))
        prm = np.zeros((len(Is), 24), dtype=np.float32)
        for r, p in enumerate(params[c0:c0 + CH]):
            if isinstance(p, torch.Tensor):
                p = p.detach().cpu().numpy()
            prm[r, :p.shape[1]] = p[0]                              # float64 -> float32, as torch.tensor([param], dtype=torch.float)
        prm = torch.from_numpy(prm)
        row_ops = [[int(o)] for o in ops[c0:c0 + CH]]
        ops_dev, ops_host = TF._prep_row_ops(row_ops, len(Is), dev)
        mask, mask_ch = None, 0
        if mask_list is not None and any(mk is not None for mk in mask_list[c0:c0 + CH]):
            mask_ch = max(mk.shape[1] for mk in mask_list[c0:c0 + CH] if mk is not None)
            ones = torch.ones(1, mask_ch, img.shape[2], img.shape[3], device=dev)
            mask = torch.cat([ones if mk is None else mk.to(dev).float().expand(1, mask_ch, -1, -1)
                              for mk in mask_list[c0:c0 + CH]], 0).contiguous()
        out, l1 = TF._rows_forward_raw(ops_dev, ops_host, img, mask, mask_ch, prm.to(dev), tgt, True, True, L)
        vals += (l1 / numel).tolist()
        outs += list(out.split(1))                                  # (views of the step's output tensor, one per row)
    return outs, vals


def _fit_sharded(states, I_gt, problems, executor, state_pair, counter, numel, group):
    """Candidate-sharded fitting: every rank holds the same states and the same problem
    list; rank r runs the Nelder-Mead fits of problems r, r + R, ... and ONE all_gather of a (n, nfev, x[24]) float64 record
    per fit gives every rank the whole table.  A fit does not depend on which other fits share its launches, so the
    table -- and everything selected from it -- is the same for every rank count."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = list(range(rank, len(problems), world))
    per = (len(problems) + world - 1) // world
    local = torch.zeros(per, 26, dtype=torch.float64)
    fits = fit_params_nelder_mead(states, I_gt, [problems[k][:2] for k in mine], executor, state_target=state_pair, numel=numel)
    for slot, r in enumerate(fits):
        local[slot, 0], local[slot, 1] = len(r.x), r.nfev
        local[slot, 2:2 + len(r.x)] = torch.from_numpy(np.asarray(r.x, dtype=np.float64))
    local = local.to(states.device)
    table = torch.empty(world * per, 26, dtype=torch.float64, device=states.device)
    dist.all_gather_into_tensor(table, local, group=group)
    table = table.cpu().view(world, per, 26)
    params, nfevs = [], []
    for k in range(len(problems)):
        row = table[k % world, k // world]
        n = int(row[0])
        params.append(row[2:2 + n].numpy().copy()[None, :])
        nfevs.append(int(row[1]))
    if counter is not None:
        counter[0] += sum(nfevs)
    return params, nfevs


def _step_minima_sharded(dists, problems, live, device_, group):
    """The lowest candidate distance of every live pair, agreed by ONE all_reduce(MIN) over packed (distance, candidate)
    keys (dist.best_candidate): rank r contributes the candidates it fitted.  -> {pair: (min dist, problem index)}"""
    import torch.distributed as dist
    from . import dist as D
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    row = {m: r for r, m in enumerate(live)}
    width = max(1, max(sum(1 for k in range(rank0, len(problems), world)) for rank0 in range(world)))
    sc = torch.full((len(live), width), float('inf'))
    ids = torch.full((len(live), width), 0x7FFFFFFF, dtype=torch.int64)
    fill = [0] * len(live)
    for k in range(rank, len(problems), world):
        r = row[problems[k][2]]
        sc[r, fill[r]], ids[r, fill[r]] = dists[k], k
        fill[r] += 1
    best, bid = D.best_candidate(sc.to(device_), ids.to(device_), group=group)
    best, bid = best.tolist(), bid.tolist()
    return {m: (best[row[m]], bid[row[m]]) for m in live}


def beam_search_batch(I_0, I_gt, executor, beam_size, operations, operation_names, max_step, err, dist_type='L1',
                      optimizer='Nelder-Mead', replace=False, _variant='default', _eps=0.05, counter=None, txt=None,
                      trace=None, shard_fits=False, group=None, masks=None, mask_op_idx=None, images='all'):
    """`beam_search` (utils/beam_search.py:196-264) for M image pairs at once: I_0, I_gt (M,3,H,W).

    Every pair runs the reference's beam search unchanged; what is shared is the work: all (pair, beam state,
    operator) fits of a step advance in lock-step on the device and are applied / scored by one launch.
    Returns a list of M (actions, Is) tuples, each exactly what `beam_search` returns for that pair.
    `trace`: an empty list that receives, per pair, {'steps': [{'candidates': [{'parent', 'op', 'param', 'dist', 'nfev'}],
    'sort_dists', 'sort_order'}]} -- every candidate evaluated and the array / order of the step's argsort (the
    transcript format of oracle/make_planner_golden_full.py).
    `masks` / `mask_op_idx` (GIER, preprocess/gen_greedy_seqs_GIER.py:60-62): per pair a list of (1, 1|3, H, W) masks and the
    operator index each belongs to (< 0: the global all-ones mask, offered to every operator); operator `op` is then
    tried once per mask that is global or belongs to it, edits only inside it (Operator.execute's blend), and the
    chosen mask's position in the list is recorded as a fourth field of the action.
    `shard_fits` (inside an initialised process group, every rank calling with the SAME pairs): candidate-sharded mode --
    the fits of a step are split over the ranks, the table of fitted parameters is all-gathered and the step's best
    candidate per pair is agreed with one all_reduce(MIN) (NCCL on GPUs); the result is the same for every rank count.
    `images`: which edited images come back on the host -- 'all' (every beam's, as the reference returns them), 'top' (the
    best sequence's only, which is what the dataset driver writes to disk, preprocess/gen_greedy_seqs_FiveK.py:86-88; the
    other beams' image lists are empty) or 'none'."""
    assert dist_type == 'L1', 'only the L1 distance is implemented'
    assert images in ('all', 'top', 'none')
    if shard_fits:
        import torch.distributed as _dist
        shard_fits = _dist.is_available() and _dist.is_initialized() and _dist.get_world_size(group) > 1
        assert not shard_fits or optimizer == 'Nelder-Mead', 'candidate sharding covers the Nelder-Mead fits'
    I_0 = I_0.to(device) if not I_0.is_cuda else I_0
    I_gt = I_gt.to(I_0.device)
    M = I_0.shape[0]
    numel = float(I_gt[0:1].numel())
    st = [{'min_dist': float('inf'), 'sequences': [[[], float('inf')]], 'I_buff': [I_0[m:m + 1]], 'alive': True}
          for m in range(M)]
    # GIER masks: one device tensor of all local masks; the global (all-ones) mask is "no mask"
    mask_bank, mask_gidx = [], None
    if masks is not None:
        assert len(masks) == M and len(mask_op_idx) == M, 'one list of masks (and of operator indices) per pair'
        assert not shard_fits, 'candidate sharding does not carry masks'
        mask_gidx = []
        for m in range(M):
            assert len(masks[m]) == len(mask_op_idx[m])
            idx = []
            for mk, oi in zip(masks[m], mask_op_idx[m]):
                if oi < 0:
                    idx.append(-1)
                else:
                    idx.append(len(mask_bank))
                    mask_bank.append(mk.to(I_0.device).float())
            mask_gidx.append(idx)
        chans = max([mk.shape[1] for mk in mask_bank] + [1])
        masks_dev = torch.cat([mk.expand(1, chans, -1, -1) for mk in mask_bank], 0).contiguous() if mask_bank else None
    if trace is not None:
        trace.extend({'steps': []} for _ in range(M))
    for i in range(max_step):
        live = [m for m in range(M) if st[m]['alive']]
        if not live:
            break
        # -- every (pair, beam state, operator) of this step (utils/beam_search.py:220-223)
        problems, states, state_pair = [], [], []       # problem: (state index, operation, pair, beam index)
        for m in live:
            seqs, buff = st[m]['sequences'], st[m]['I_buff']
            for j in range(len(buff)):
                s_idx = len(states)
                states.append(buff[j])
                state_pair.append(m)
                step_ops = [operations[i]] if _variant == 'fixed_order' else operations
                for operation in step_ops:
                    if not replace and operation in [operation_names.index(v[0]) for v in seqs[j][0]]:
                        continue
                    if masks is None:
                        problems.append((s_idx, operation, m, j, None))
                    else:
                        for k, oi in enumerate(mask_op_idx[m]):
                            if oi < 0 or oi == operation:
                                problems.append((s_idx, operation, m, j, k))
        # -- fit all of them (utils/beam_search.py:229)
        states_cat = torch.cat(states, 0).contiguous() if problems else None
        if shard_fits and problems:
            params, nfevs = _fit_sharded(states_cat, I_gt, problems, executor, state_pair, counter,
                                         numel, group)
        elif optimizer == 'Nelder-Mead' and problems:
            use_masks = masks is not None and masks_dev is not None
            fits = fit_params_nelder_mead(states_cat, I_gt, [(s, op) for s, op, _, _, _ in problems],
                                          executor, state_target=state_pair, counter=counter, numel=numel,
                                          masks=masks_dev if use_masks else None,
                                          prob_mask=[mask_gidx[m][k] for _, _, m, _, k in problems] if use_masks else None)
            params = [np.asarray(r.x, dtype=np.float64)[None, :] for r in fits]     # (1, n) float64, as the reference's tensors
            nfevs = [r.nfev for r in fits]
        else:
            assert masks is None, 'masks are carried by the Nelder-Mead fits'
            params = [get_param(states[s], I_gt[m:m + 1], txt, op, executor, None, dist_type, optimizer)[0]
                      for s, op, m, _, _ in problems]
            nfevs = [0] * len(problems)
        # -- apply + score (utils/beam_search.py:230-237)
        mask_list = None
        if masks is not None:
            mask_list = [None if mask_gidx[m][k] < 0 else masks_dev[mask_gidx[m][k]:mask_gidx[m][k] + 1] for _, _, m, _, k in problems]
        outs, dists = _score_outputs([states[s] for s, _, _, _, _ in problems], [op for _, op, _, _, _ in problems], params,
                                     [I_gt[m:m + 1] for _, _, m, _, _ in problems], executor, mask_list,
                                     gather=None if not problems else (states_cat, [s for s, _, _, _, _ in problems],
                                                                       I_gt if I_gt.dtype == torch.float32 else I_gt.float(),
                                                                       [m for _, _, m, _, _ in problems]))
        # -- the reference's bookkeeping, pair by pair (utils/beam_search.py:239-259)
        minima = _step_minima_sharded(dists, problems, live, I_0.device, group) if shard_fits and problems else None
        by_pair = {m: [] for m in live}
        for k, (s, op, m, j, mk) in enumerate(problems):
            by_pair[m].append((j, op, params[k], outs[k], dists[k], nfevs[k], mk))
        for m in live:
            S = st[m]
            all_candidates, I_tmp_list, tmp_min_dists = [], [], []
            no_update_flag, finish_flag = True, False
            for j, operation, param, I_out, dist, _, mk in by_pair[m]:
                if _variant == 'eps_greedy' or dist < S['min_dist']:
                    tmp_min_dists.append(dist)
                    act = (operation_names[operation], param[0].tolist(), dist) + (() if mk is None else (mk,)) + (I_out,)
                    candidate = [S['sequences'][j][0] + [act], dist]
                    all_candidates.append(candidate)
                    I_tmp_list.append(I_out)
                    if _variant != 'eps_greedy':
                        no_update_flag = False
                    if dist < err:
                        finish_flag = True
            S['min_dist'] = min(tmp_min_dists) if len(tmp_min_dists) > 0 else S['min_dist']
            if minima is not None and by_pair[m]:
                # candidate-sharded mode: the all-reduced minimum is the step's new min_dist (utils/beam_search.py:250)
                agreed = minima[m][0] if _variant == 'eps_greedy' else min(S['min_dist'], minima[m][0])
                assert agreed == S['min_dist'], ('ranks disagree on the best candidate', m, agreed, S['min_dist'])
                S['min_dist'] = agreed
            if len(all_candidates) < beam_size:
                all_candidates += S['sequences']
                I_tmp_list += S['I_buff']
            dists_arr = np.array([v[1] for v in all_candidates])
            order = np.argsort(dists_arr)
            buf_idx = [int(v) for v in order[:beam_size]]
            if _variant == 'eps_greedy' and _eps_rng.random() < _eps:
                # utils/beam_search_eps_greedy.py:299-302: the SEQUENCES are drawn at random (random.choices draws
                # the same floor(random() * n) indices for a list and for a range), the image buffer stays the sorted one
                seq_idx = _eps_rng.choices(range(len(all_candidates)), k=beam_size)
            else:
                seq_idx = buf_idx
            if trace is not None:
                trace[m]['steps'].append({
                    'candidates': [{'parent': j, 'op': int(op), 'param': [float(v) for v in np.asarray(p).reshape(-1)],
                                    'dist': float(d), 'nfev': int(nf), 'mask': mk} for j, op, p, _, d, nf, mk in by_pair[m]],
                    'sort_dists': [float(v) for v in dists_arr], 'sort_order': [int(v) for v in order]})
            # the kept images are rows of this step's output tensor: copy them out (once) so that it can be freed
            clones = {}

            def own(idx):
                if idx not in clones:
                    img = I_tmp_list[idx]
                    clones[idx] = img.clone() if img._base is not None else img
                return clones[idx]
            seqs = []
            for idx in seq_idx:
                seq = all_candidates[idx]
                if seq[0] and seq[0][-1][-1] is I_tmp_list[idx]:
                    seq = [seq[0][:-1] + [seq[0][-1][:-1] + (own(idx),)], seq[1]]
                seqs.append(seq)
            S['sequences'], S['I_buff'] = seqs, [own(idx) for idx in buf_idx]
            if no_update_flag or finish_flag:
                S['alive'] = False
    # the reference returns CPU images (:236): every distinct kept image once, in a few large device -> host copies
    # (sequences of a pair share their prefixes; ~1 400 separate .cpu() calls cost a 64-pair search 0.14 s)
    uniq, order = {}, []
    for m in range(M):
        for b, seq in enumerate(st[m]['sequences']):
            if images == 'none' or (images == 'top' and b > 0):
                continue
            for act in seq[0]:
                if id(act[-1]) not in uniq:
                    uniq[id(act[-1])] = len(order)
                    order.append(act[-1])
    host = []
    CH = 512
    for c0 in range(0, len(order), CH):
        host.append(torch.cat(order[c0:c0 + CH], 0).cpu())
    results = []
    for m in range(M):
        seqs = st[m]['sequences']
        actions = [[act[:-1] for act in seq[0]] for seq in seqs]
        Is = [[] if images == 'none' or (images == 'top' and b > 0) else
              [host[uniq[id(act[-1])] // CH][uniq[id(act[-1])] % CH:uniq[id(act[-1])] % CH + 1] for act in seq[0]]
              for b, seq in enumerate(seqs)]
        results.append((actions, Is))
    return results


_PIPE_LOCAL = threading.local()      # per worker thread: its CUDA stream
_PIPE_POOLS = {}                     # workers -> the persistent thread pool


def beam_search_pipelined(batches, executor, beam_size, operations, operation_names, max_step, err, workers=2, counter=None,
                          **kwargs):
    """`beam_search_batch` over a stream of batches -- the dataset loop of preprocess/gen_greedy_seqs_FiveK.py:44-64 -- with
    `workers` batches in flight: each runs on its own thread and CUDA stream, so the host bookkeeping of one batch (selection,
    records, result copies: about half of a batch's wall time) overlaps the device fits of another, and the resident
    Nelder-Mead launches of two batches fill each other's tails.  Batches are independent, so every result equals the
    sequential call's.  batches: an iterable of (I_0, I_gt) CUDA tensor pairs, or (I_0, I_gt, kwargs) with keyword arguments
    of that batch alone (its masks / mask_op_idx) -- consumed at most `workers` ahead; returns the list of beam_search_batch
    results in order.  (Not for the eps-greedy variant, whose random draws are ordered.)"""
    import concurrent.futures as cf
    assert kwargs.get('_variant', 'default') != 'eps_greedy', 'eps-greedy draws from one ordered random stream'
    if workers <= 1:
        return [beam_search_batch(bt[0], bt[1], executor, beam_size, operations, operation_names, max_step, err, counter=counter,
                                  **dict(kwargs, **(bt[2] if len(bt) > 2 else {}))) for bt in batches]
    results, counts, pending = {}, {}, []
    local = _PIPE_LOCAL

    def work(k, I_0, I_gt, ready, kw):
        dev = I_0.device
        torch.cuda.set_device(dev)                                  # (a new thread starts on device 0)
        if not hasattr(local, 'streams'):
            local.streams = {}
        stream = local.streams.get(dev)
        if stream is None:
            stream = local.streams[dev] = torch.cuda.Stream(dev)
        cnt = [0]
        with torch.cuda.stream(stream):
            stream.wait_event(ready)                                # the batch was produced on the caller's stream
            res = beam_search_batch(I_0, I_gt, executor, beam_size, operations, operation_names, max_step, err, counter=cnt, **kw)
            stream.synchronize()
        I_0.record_stream(stream)
        I_gt.record_stream(stream)
        return k, res, cnt[0]

    # the worker threads (and with them their CUDA streams, the allocator's pools of those streams and the library's
    # workspaces) live as long as the process: a fresh pool per call meant fresh streams, and a second or two of cudaMalloc
    # at the start of every call
    pool = _PIPE_POOLS.get(workers)
    if pool is None:
        pool = _PIPE_POOLS[workers] = cf.ThreadPoolExecutor(max_workers=workers, thread_name_prefix='t2o-planner')

    def drain(fut):
        k, res, c = fut.result()
        results[k], counts[k] = res, c
    try:
        for k, bt in enumerate(batches):
            I_0, I_gt = bt[0], bt[1]
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(I_0.device))
            pending.append(pool.submit(work, k, I_0, I_gt, ready, dict(kwargs, **(bt[2] if len(bt) > 2 else {}))))
            if len(pending) >= workers + 1:
                drain(pending.pop(0))
        while pending:
            drain(pending.pop(0))
    finally:
        for fut in pending:                                         # (an exception above: let the batches in flight finish)
            try:
                fut.result()
            except Exception:
                pass
    if counter is not None:
        counter[0] += sum(counts.values())
    return [results[k] for k in range(len(results))]


def beam_search(I_0, I_gt, txt, executor, discriminator, beam_size, operations, operation_names, max_step, err,
                dist_type, optimizer, replace=False, _variant='default', _eps=0.05, counter=None, shard_fits=False, group=None):
    """utils/beam_search.py:196-264 -- same arguments and return value:
    actions: list(beam) of list(step) of (op_name, param list, dist); Is: same nesting of CPU images."""
    return beam_search_batch(I_0, I_gt, executor, beam_size, operations, operation_names, max_step, err, dist_type,
                             optimizer, replace, _variant, _eps, counter, txt, shard_fits=shard_fits, group=group)[0]


def beam_search_gier(I_0, I_gt, txt, mask, mask_op_idx, executor, beam_size, operations, operation_names, max_step, err, dist_type,
                     optimizer, replace=False):
    """The call the GIER planner driver makes (preprocess/gen_greedy_seqs_GIER.py:71):
    beam_search(input, target, req_idx, mask, mask_op_idx, executor, beam_size, operations, ...), with `mask` the list
    [global all-ones mask] + local masks and `mask_op_idx` = [-1] + the operator each local mask belongs to.  The
    reference's own beam_search does not take these arguments (the driver does not run as committed); here every operator
    is tried once per mask that is global or belongs to it.  Actions: (op_name, param list, dist, position in `mask`)."""
    return beam_search_batch(I_0, I_gt, executor, beam_size, operations, operation_names, max_step, err, dist_type, optimizer,
                             replace, txt=txt, masks=[list(mask)], mask_op_idx=[list(mask_op_idx)])[0]


def beam_search_fixed_order(I_0, I_gt, txt, executor, beam_size, operations, operation_names, max_step, err, dist_type,
                            optimizer, replace=False):
    """utils/beam_search_fixed_order.py:225-293 (no discriminator argument; one operator per step)."""
    return beam_search(I_0, I_gt, txt, executor, None, beam_size, operations, operation_names, max_step, err, dist_type,
                       optimizer, replace, _variant='fixed_order')


def beam_search_eps_greedy(I_0, I_gt, txt, executor, discriminator, beam_size, operations, operation_names, max_step, err,
                           dist_type, optimizer, eps=0.05, replace=False):
    """utils/beam_search_eps_greedy.py:238-309"""
    return beam_search(I_0, I_gt, txt, executor, discriminator, beam_size, operations, operation_names, max_step, err,
                       dist_type, optimizer, replace, _variant='eps_greedy', _eps=eps)
