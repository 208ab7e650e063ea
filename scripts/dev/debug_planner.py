"""GPU debug probe: planner transcripts vs the reference goldens, with evidence for every failing pair."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import t2onet_b200 as T
import t2onet_b200.functional as TF
from oracle import ops as O
from planner_compare import compare_runs, NAMES
G = os.path.join(ROOT, 'tests', 'golden')

def load(mode):
    rec = json.load(open(os.path.join(G, 'planner_full_%s.json' % mode)))
    d = np.load(os.path.join(G, rec['settings'].get('images', 'planner_full_%s.npz' % mode)))
    return rec, (torch.from_numpy(d['I0']).float() / 255).cuda(), (torch.from_numpy(d['Igt']).float() / 255).cuda()

def eval_fn(ex, I0_m, Igt_m):
    def fn(parent_actions, op, xs):
        img = I0_m
        for pop, pparam in parent_actions:
            img = T.planner.execute(img, pop, torch.tensor([pparam], device='cuda', dtype=torch.float32), ex)
        prm = torch.zeros(len(xs), 24)
        prm[:, 0] = torch.tensor(xs, dtype=torch.float64).float()
        l1 = TF.score_candidates(img, Igt_m, [0] * len(xs), [op] * len(xs), prm)
        return (l1 / float(img.numel())).tolist()
    return fn

for mode in sys.argv[1:] or ('c5', 'c3'):
    rec, I0, Igt = load(mode)
    st = rec['settings']
    ex = T.Executor(T.default_options()).cuda()
    trace = []
    res = T.planner.beam_search_batch(I0, Igt, ex, st['beam'], st['operations'], O.ACTION_NAMES, st['max_step'], st['err'], trace=trace)
    for m, pair in enumerate(rec['pairs']):
        try:
            v, detail = compare_runs(pair['steps'], trace[m]['steps'], st['beam'], st['err'], 5e-4, 1e-4, 2e-3,
                                     eval_fn=eval_fn(ex, I0[m:m + 1], Igt[m:m + 1]), eval_tol=2e-6)
            print(mode, m, v, str(detail)[:200])
        except AssertionError as e:
            print(mode, m, 'FAIL', str(e)[:300])
            a = e.args[0]
            if isinstance(a, tuple) and len(a) >= 3 and isinstance(a[2], tuple):
                s, key = a[1], a[2]
                rc = [c for c in pair['steps'][s]['candidates']]
                # reference candidate and ours
                from planner_compare import replay_selection
                rr = replay_selection(pair['steps'], st['beam'], st['err'])[s]
                gg = replay_selection(trace[m]['steps'], st['beam'], st['err'])[s]
                print('    ref dist %.9f ours %.9f' % (rr['cands'][key], gg['cands'][key]))
                for c in pair['steps'][s]['candidates']:
                    pseq = rr['beam_in'][c['parent']]
                    if (pseq, c['op']) == key:
                        print('    ref  param', c['param'][:4], 'nfev', c['nfev'])
                        hist = c.get('hist')
                for c in trace[m]['steps'][s]['candidates']:
                    pseq = gg['beam_in'][c['parent']]
                    if (pseq, c['op']) == key:
                        print('    ours param', c['param'][:4], 'nfev', c['nfev'])
                        ours_c = c
                # parent distances
                if key[0]:
                    pk = (key[0][:-1], NAMES.index(key[0][-1]))
                    rp = replay_selection(pair['steps'], st['beam'], st['err'])[s - 1]['cands'].get(pk)
                    gp = replay_selection(trace[m]['steps'], st['beam'], st['err'])[s - 1]['cands'].get(pk)
                    print('    parent dist ref %s ours %s' % (rp, gp))
                if key[0]:
                    # the two parent states (reference's recorded parent actions / ours) and the two fits, cross-evaluated
                    acts_r, acts_g = rr['acts'][key[0]], gg['acts'][key[0]]
                    for tag, acts in (('ref-parent', acts_r), ('our-parent', acts_g)):
                        img = I0[m:m + 1]
                        for pop, pparam in acts:
                            img = T.planner.execute(img, pop, torch.tensor([pparam], device='cuda', dtype=torch.float32), ex)
                        d0 = T.planner.get_dist(img, Igt[m:m + 1]).item()
                        prm = torch.zeros(401, 24)
                        grid = torch.linspace(-1, 3, 401)
                        prm[:, 0] = grid
                        l1 = TF.score_candidates(img, Igt[m:m + 1], [0] * 401, [key[1]] * 401, prm) / float(img.numel())
                        k = int(l1.argmin())
                        print('    %s: dist %.6f; %s sweep min %.6f at p=%.3f; at p=0: %.6f' % (tag, d0, NAMES[key[1]], l1[k].item(), grid[k].item(), l1[100].item()))
                if hist:
                    xs = [h[0] for h in hist]
                    f_sc = eval_fn(ex, I0[m:m + 1], Igt[m:m + 1])(rr['acts'][key[0]], key[1], xs)
                    # the same through Executor.execute + get_dist, and the oracle in fp32 / fp64 on the GPU
                    img = I0[m:m + 1]
                    for pop, pparam in rr['acts'][key[0]]:
                        img = T.planner.execute(img, pop, torch.tensor([pparam], device='cuda', dtype=torch.float32), ex)
                    for x, fr, fs in list(zip(xs, [h[1] for h in hist], f_sc))[:8]:
                        out = T.planner.execute(img, key[1], torch.tensor([[x]], device='cuda', dtype=torch.float32), ex)
                        fe = T.planner.get_dist(out, Igt[m:m + 1]).item()
                        o32 = O.execute(key[1], img.cpu(), torch.tensor([[x]], dtype=torch.float32))
                        f32 = ((o32 - Igt[m:m + 1].cpu()).norm(1) / o32.numel()).item()
                        try:
                            o64 = O.execute(key[1], img.cpu().double(), torch.tensor([[x]], dtype=torch.float64))
                            f64 = ((o64 - Igt[m:m + 1].cpu().double()).norm(1) / o64.numel()).item()
                        except Exception:
                            f64 = ((o32.double() - Igt[m:m + 1].cpu().double()).norm(1) / o32.numel()).item()
                        print('      x %+.7f ref %.9f scorer %.9f execute %.9f oracle32 %.9f oracle64 %.9f' % (x, fr, fs, fe, f32, f64))
