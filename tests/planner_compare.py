"""Compare a planner run with a transcript recorded from the reference (test infrastructure).

Transcript format (oracle/make_planner_golden_full.py; planner.beam_search_batch(trace=...) and the oracle planner emit
the same): per step, every candidate the planner evaluated -- {'parent' (beam index), 'op', 'param', 'dist'} in
evaluation order -- plus 'sort_dists' / 'sort_order', the array handed to np.argsort and its result.

The north star's rule: the chosen action sequences must be identical "whenever no candidate scores are tied within
tolerance".  This module makes the exception checkable instead of asserted: a step of the run under test may keep a
different beam than the reference only if the REFERENCE'S OWN recorded distances of the two competing candidates (or of
the candidate and the threshold it was compared with) differ by at most `tie_tol`.  From the first tolerated
divergence on the two runs explore different states, so the comparison of that pair stops there (its end result is
still bounded: final distance within `tie_tol` of the reference's).

The same rule one level down, inside a parameter fit: Nelder-Mead's path is a function of the ORDER of the values it
evaluates.  If a 1-parameter fit of the run under test ends elsewhere than the reference's (beyond the fit tolerance),
the reference's recorded evaluation history (x_i, f_i) is re-scored by the implementation under test (`eval_fn`): every
value must agree with the reference's within `eval_tol` (that is the parity statement for the scores themselves), and
there must be a pair of evaluations whose reference values differ by at most 2 eval_tol and whose order the re-scored
values reverse -- two candidate scores tied within the L1's own rounding noise that decided the path.  Otherwise the
mismatch is an error."""
NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
CURVE_OPS = (3, 5)
N_PARAMS = {3: 24, 5: 8}            # parameters of the curve operators (the others have one): Nelder-Mead's maxfev is 200 N


def replay_selection(steps, beam, err, variant='default'):
    """Re-run the reference's bookkeeping (utils/beam_search.py:239-259) over a transcript.
    Returns per step: {'beam_in': [op-name tuples], 'cands': {(parent seq, op): dist}, 'all': [(seq, dist)] the list
    that was sorted, 'beam_out': [(seq, dist)], 'min_dist_in': float, 'finish': bool, 'no_update': bool}."""
    sequences = [((), float('inf'))]
    min_dist = float('inf')
    out = []
    acts = {(): []}                 # op-name sequence -> [(op, param)] that produced it
    unconv = {(): False}            # op-name sequence -> some fit on the way stopped at maxfev (200 N evaluations)
    for st in steps:
        kept, cands = [], {}
        finish, no_update = False, True
        hist = {}
        for c in st['candidates']:
            pseq = sequences[c['parent']][0]
            seq = pseq + (NAMES[c['op']],)
            cands[(pseq, c['op'])] = c['dist']
            hist[(pseq, c['op'])] = c.get('hist')
            acts[seq] = acts[pseq] + [(c['op'], c['param'])]
            unconv[seq] = unconv[pseq] or c.get('nfev', 0) >= 200 * N_PARAMS.get(c['op'], 1)
            if variant == 'eps_greedy' or c['dist'] < min_dist:
                kept.append((seq, c['dist']))
                if variant != 'eps_greedy':
                    no_update = False
                if c['dist'] < err:
                    finish = True
        rec = {'beam_in': [s for s, _ in sequences], 'cands': cands, 'min_dist_in': min_dist, 'hist': hist,
               'acts': dict(acts), 'unconv': dict(unconv)}
        if kept:
            min_dist = min(d for _, d in kept)
        all_c = kept + (sequences if len(kept) < beam else [])
        assert len(all_c) == len(st['sort_dists']), 'transcript inconsistent with the selection rule'
        for (s, d), d2 in zip(all_c, st['sort_dists']):
            assert d == d2 or (d != d and d2 != d2), 'transcript inconsistent with the selection rule'
        sequences = [all_c[i] for i in st['sort_order']][:beam]
        rec.update({'all': all_c, 'beam_out': sequences, 'finish': finish, 'no_update': no_update})
        out.append(rec)
    return out


def fit_level_tie(hist, f_got, eval_tol):
    """hist: the reference's [(x, f)] of one fit; f_got: the same x's scored by the implementation under test.
    -> (i, j) of a pair tied within 2 eval_tol in the reference whose order f_got reverses, or None."""
    n = len(hist)
    for i in range(n):
        for j in range(i + 1, n):
            a, b = hist[i][1], hist[j][1]
            if abs(a - b) <= 2 * eval_tol and ((a < b) != (f_got[i] < f_got[j]) or (a == b) != (f_got[i] == f_got[j])):
                return i, j
    return None


def compare_runs(ref_steps, got_steps, beam, err, tie_tol, fit_tol_scalar, fit_tol_curve, variant='default',
                 eval_fn=None, eval_tol=2e-6, ref_noise_cap=None, unconv_band=None):
    """-> (verdict, detail).  verdict: 'exact' (same candidates within the fit tolerances, identical beams at every
    step, same number of steps), 'tie' (first divergence justified by the reference's own distances; detail says
    where), 'path' (as 'tie', but the justification needs the path dependence of a fit that stopped at maxfev, see
    below), or raises AssertionError with the evidence.

    `ref_noise_cap`: the reference's recorded evaluations may carry their own summation noise (torch's CPU fp32
    norm(1), see oracle/make_planner_golden_full.py): the noise of a fit is MEASURED as max |f_ref - f_rescored| over its
    history (must be <= the cap) and replaces eval_tol in the fit-level rule.
    `unconv_band`: Nelder-Mead stops the 24- / 8-parameter fits at maxfev = 200 N, far from convergence; where it stands
    then depends on the last bits of every value it has seen, in the reference as much as here.  A candidate whose own
    fit, or a fit that produced its parent state, hit maxfev (in either run) is compared within `unconv_band` only, and
    its measured discrepancy |dist - ref dist| counts as the uncertainty of its score when two candidates swap places."""
    R = replay_selection(ref_steps, beam, err, variant)
    G = replay_selection(got_steps, beam, err, variant)
    worst = {'scalar': 0.0, 'curve': 0.0, 'unconv': 0.0}
    for s in range(max(len(R), len(G))):
        if s >= len(R) or s >= len(G):
            # one run stopped earlier: finish_flag / no_update_flag disagreed at step s-1
            r, g = R[s - 1], G[s - 1]
            slack = tie_tol + (worst['unconv'] if unconv_band is not None else 0.0)
            near_err = any(abs(d - err) <= slack for d in r['cands'].values())
            near_min = any(abs(d - r['min_dist_in']) <= slack for d in r['cands'].values())
            assert near_err or near_min, ('step count differs without a tie', s, len(R), len(G))
            return ('tie' if slack == tie_tol else 'path'), 'stopped at step %d vs %d: a candidate within %.0e of err / of the previous minimum' % (len(G), len(R), slack)
        r, g = R[s], G[s]
        assert r['beam_in'] == g['beam_in'], ('beams differ entering step %d' % s, r['beam_in'], g['beam_in'])
        assert set(r['cands']) == set(g['cands']), ('different candidates evaluated at step %d' % s)
        disc = {}
        for key, d_ref in r['cands'].items():
            seq = key[0] + (NAMES[key[1]],)
            curve = key[1] in CURVE_OPS or any(NAMES.index(nm) in CURVE_OPS for nm in key[0])
            diff = abs(g['cands'][key] - d_ref)
            # (with a noisy reference every 8- / 24-parameter fit is path dependent: it ends where the noise stops it)
            if unconv_band is not None and (r['unconv'][seq] or g['unconv'][seq] or (ref_noise_cap is not None and curve)):
                assert diff <= unconv_band, ('candidate behind an unconverged fit off by more than the band', s, key, g['cands'][key], d_ref)
                disc[seq] = diff
                worst['unconv'] = max(worst['unconv'], diff)
                continue
            kind = 'curve' if curve else 'scalar'
            tol = fit_tol_curve if curve else fit_tol_scalar
            if diff > tol and not curve and eval_fn is not None and r['hist'].get(key):
                # a 1-parameter fit that ended elsewhere: re-score the reference's own evaluation history
                hist = r['hist'][key]
                f_got = eval_fn(r['acts'][key[0]], key[1], [h[0] for h in hist])
                worst_eval = max(abs(a - h[1]) for a, h in zip(f_got, hist))
                depth_tol = eval_tol * (1 + 4 * len(key[0]))      # deeper states carry the earlier fits' 1e-4 parameter slack
                if ref_noise_cap is not None:
                    assert worst_eval <= ref_noise_cap, ('re-scored evaluations beyond the reference noise cap', s, key, worst_eval)
                    depth_tol = max(depth_tol, worst_eval)
                else:
                    assert worst_eval <= depth_tol, ('re-scored evaluations off', s, key, worst_eval)
                pair = fit_level_tie(hist, f_got, depth_tol)
                assert pair is not None, ('fit ended elsewhere without a tied pair of evaluations', s, key, g['cands'][key], d_ref)
                i, j = pair
                return 'tie', ('step %d, %s after %s: Nelder-Mead path decided by evaluations %d / %d: reference f = %.9f / %.9f '
                               '(|diff| %.1e%s), re-scored %.9f / %.9f; the fit ends at %.6f vs the reference\'s %.6f' % (
                                   s, NAMES[key[1]], '>'.join(key[0]) or 'the input', i, j, hist[i][1], hist[j][1],
                                   abs(hist[i][1] - hist[j][1]),
                                   '' if ref_noise_cap is None else '; measured noise of the reference\'s values %.1e' % worst_eval,
                                   f_got[i], f_got[j], g['cands'][key], d_ref))
            worst[kind] = max(worst[kind], diff)
            assert diff <= tol, ('candidate distance off', s, key, g['cands'][key], d_ref)
        rb, gb = [q for q, _ in r['beam_out']], [q for q, _ in g['beam_out']]
        if rb != gb:
            # every position where the kept sequences differ must be a tie IN THE REFERENCE'S OWN numbers
            ref_d = {q: d for q, d in r['all']}
            ref_d.update({pseq + (NAMES[op],): d for (pseq, op), d in r['cands'].items() if pseq + (NAMES[op],) not in ref_d})
            ev, used_disc = [], False
            for k in range(max(len(rb), len(gb))):
                a = rb[k] if k < len(rb) else None
                b = gb[k] if k < len(gb) else None
                if a == b:
                    continue
                da = ref_d.get(a, r['min_dist_in']) if a is not None else r['min_dist_in']
                db = ref_d.get(b, r['min_dist_in']) if b is not None else r['min_dist_in']
                slack = tie_tol + disc.get(a, 0.0) + disc.get(b, 0.0)
                if abs(da - db) > tie_tol and (disc.get(a, 0.0) + disc.get(b, 0.0)) > 0.0:
                    # a candidate that left the beam may also have been pushed out by an unconverged one further up
                    slack = tie_tol + 2 * max(disc.values())
                assert abs(da - db) <= slack, ('beam differs at step %d position %d without a tie in the reference' % (s, k),
                                               a, da, b, db, slack)
                used_disc = used_disc or abs(da - db) > tie_tol
                ev.append('pos %d: ref %s (%.6f) vs got %s (ref dist %.6f)' % (k, '>'.join(a or ()), da, '>'.join(b or ()), db))
            return ('path' if used_disc else 'tie'), 'step %d: %s%s' % (s, '; '.join(ev), (
                '; unconverged-fit discrepancies up to %.1e' % max(disc.values())) if used_disc else '')
        if r['finish'] != g['finish'] or r['no_update'] != g['no_update']:
            slack = tie_tol + (max(disc.values()) if disc else 0.0)
            near_err = any(abs(d - err) <= slack for d in r['cands'].values())
            near_min = any(abs(d - r['min_dist_in']) <= slack for d in r['cands'].values())
            assert near_err or near_min, ('termination differs without a tie', s)
            return ('tie' if slack == tie_tol else 'path'), 'step %d: termination flag flipped by a candidate within %.0e of err / the previous minimum' % (s, slack)
    return 'exact', 'max |dist - ref|: converged scalar fits %.1e, converged curve fits %.1e, behind unconverged fits %.1e' % (
        worst['scalar'], worst['curve'], worst['unconv'])
