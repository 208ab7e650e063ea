"""Operation-planner oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates utils/beam_search.py of the reference (L1 distance branch; the
discriminator branches reference undefined names and are dead code, SURVEY.md
section 2 row 4) on top of ``oracle.ops``.  The per-(state, operator) parameter
fit is scipy's Nelder-Mead exactly as the reference calls it
(``minimize(func, param0, method='Nelder-Mead')``, utils/beam_search.py:88; scipy
is a third-party dependency, unpinned in the reference, 1.18.1 in this image), or
torch Adam / L-BFGS as in ``gd_minimize`` (utils/beam_search.py:94-128).
"""
import random

import numpy as np
import torch
from scipy.optimize import minimize

from . import ops as O


def get_dist(x1, x2, dist_type='L1'):
    """utils/beam_search.py:170-180 (L1 only)."""
    assert dist_type == 'L1', '{} is invalid distance'.format(dist_type)
    return O.l1_dist(x1, x2)


def execute(I, operation, param, executor):
    """utils/beam_search.py:165-167"""
    return executor.execute(I, operation, None, features=None, specified_param=param, has_noise=False)[0]


def get_param_naive(img, out, param0, executor, op_ind, counter=None, mask=None):
    """utils/beam_search.py:65-91 -- Nelder-Mead over the operator parameters.  `mask`: the argument the reference's
    signature carries (:65) but never hands to the executor (:79 passes None); the GIER extension below does."""
    def func(param):
        if counter is not None:
            counter[0] += 1
        p = torch.tensor(np.array([param]), dtype=torch.float)
        pred, _ = executor.execute(img, op_ind, mask, specified_param=p, has_noise=False)
        return get_dist(pred, out).item()
    res = minimize(func, param0, method='Nelder-Mead')
    return torch.tensor(np.array([list(res.x)])), res.success


def gd_minimize(func, param0, method='adam'):
    """utils/beam_search.py:94-128"""
    num_iters, tol = 1000, 1e-5
    param0.requires_grad_()
    success = False
    if method == 'lbfgs':
        success = True
        opt = torch.optim.LBFGS([param0], lr=1)

        def closure():
            opt.zero_grad()
            loss = func(param0)
            loss.backward()
            return loss
        opt.step(closure)
    elif method == 'adam':
        opt = torch.optim.Adam([param0], lr=1e-2)
        loss_prev = 10000
        for _ in range(num_iters):
            opt.zero_grad()
            loss = func(param0)
            cur = loss.item()
            if (loss_prev - cur) < tol:
                success = True
                break
            loss_prev = cur
            loss.backward()
            opt.step()
    return param0.detach(), success


def get_param(I0, I1, operation, executor, optimizer='Nelder-Mead', counter=None, mask=None):
    """utils/beam_search.py:148-162 -- zeros init for ops {0,1,2,6}, ones for {3,5}."""
    n = executor.get_param_num(operation)
    if operation in [0, 1, 2, 6]:
        param0 = torch.zeros(n)
    elif operation in [3, 5]:
        param0 = torch.ones(n)
    else:
        assert False, 'the operation is not global operation'
    if optimizer == 'Nelder-Mead':
        return get_param_naive(I0, I1, param0, executor, operation, counter, mask)
    assert mask is None
    param0 = param0.view(1, -1).repeat(I0.shape[0], 1)

    def func(p):
        if counter is not None:
            counter[0] += 1
        pred, _ = executor.execute(I0, operation, None, specified_param=p, has_noise=False)
        return get_dist(pred, I1)
    return gd_minimize(func, param0, method=optimizer)


def beam_search(I_0, I_gt, txt, executor, discriminator, beam_size, operations, operation_names, max_step,
                err, dist_type, optimizer, replace=False, variant='default', eps=0.05, counter=None, trace=None,
                mask=None, mask_op_idx=None):
    """utils/beam_search.py:196-264.

    variant='fixed_order'  -> utils/beam_search_fixed_order.py:225-293 (one operator per step)
    variant='eps_greedy'   -> utils/beam_search_eps_greedy.py:238-309 (keeps every candidate,
                              random beams with probability eps, never clears no_update_flag)
    mask / mask_op_idx: the GIER driver's arguments (preprocess/gen_greedy_seqs_GIER.py:60-71; the reference's beam_search
    does not accept them): every operator is tried once per mask that is global (index < 0) or belongs to it, the edit is
    blended inside the mask (Operator.execute), the action gets the mask's position as a fourth field.
    Returns (actions, Is) with the reference's nesting.  `trace`: a list that receives one record per step (every
    candidate evaluated + the argsort input / output), in the format of oracle/make_planner_golden_full.py.
    """
    assert dist_type == 'L1'
    min_dist = float('inf')
    sequences = [[[], float('inf')]]
    I_buff = [I_0]
    for i in range(max_step):
        all_candidates, I_tmp_list, tmp_min_dists = [], [], []
        no_update_flag, finish_flag = True, False
        step_cands = []
        for j, I in enumerate(I_buff):
            step_ops = [operations[i]] if variant == 'fixed_order' else operations
            for operation in step_ops:
                if not replace and operation in [operation_names.index(v[0]) for v in sequences[j][0]]:
                    continue
                choices = [None] if mask is None else [k for k, oi in enumerate(mask_op_idx) if oi < 0 or oi == operation]
                for k in choices:
                    mk = None if k is None or mask_op_idx[k] < 0 else mask[k]
                    param, _ = get_param(I, I_gt, operation, executor, optimizer, counter, mk)
                    I_out = executor.execute(I, operation, mk, features=None, specified_param=param, has_noise=False)[0]
                    dist = get_dist(I_out, I_gt, dist_type).item()
                    step_cands.append({'parent': j, 'op': int(operation), 'param': [float(v) for v in param[0].tolist()],
                                       'dist': float(dist), 'mask': k})
                    if variant == 'eps_greedy' or dist < min_dist:
                        tmp_min_dists.append(dist)
                        act = (operation_names[operation], param[0].tolist(), dist) + (() if k is None else (k,)) + (I_out.cpu(),)
                        cand = [sequences[j][0] + [act], dist]
                        all_candidates.append(cand)
                        I_tmp_list.append(I_out)
                        if variant != 'eps_greedy':
                            no_update_flag = False
                        if dist < err:
                            finish_flag = True
        min_dist = min(tmp_min_dists) if len(tmp_min_dists) > 0 else min_dist
        if len(all_candidates) < beam_size:
            all_candidates += sequences
            I_tmp_list += I_buff
        dists = np.array([v[1] for v in all_candidates])
        order = np.argsort(dists)
        if trace is not None:
            trace.append({'candidates': step_cands, 'sort_dists': [float(v) for v in dists],
                          'sort_order': [int(v) for v in order]})
        if variant == 'eps_greedy' and random.random() < eps:
            sequences = random.choices(all_candidates, k=beam_size)
        else:
            sequences = [all_candidates[idx] for idx in order][:beam_size]
        I_buff = [I_tmp_list[idx] for idx in order][:beam_size]
        if no_update_flag or finish_flag:
            break
    actions = [[act[:-1] for act in seq[0]] for seq in sequences]
    Is = [[act[-1] for act in seq[0]] for seq in sequences]
    return actions, Is
