"""CPU-side tests (no GPU): C-ABI surface, host logic of the binding, the coroutine Nelder-Mead against
scipy, the kernels' arithmetic header (compiled for the host) against the golden vectors, and the
multi-rank plumbing on the gloo backend (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C-ABI
def test_library_exports_every_symbol_declared_in_the_header():
    from t2onet_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 't2o.h')).read()
    declared = sorted(set(re.findall(r'^(?:int|size_t|const char \*)\s*(t2o_[a-z0-9_]+)\s*\(', header, re.M)))
    assert declared == sorted(_lib.EXPORTS)
    lib = _lib.lib()                               # builds if stale; loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.t2o_version() == 104
    assert lib.t2o_status_string(2) == b'unsupported configuration'
    assert [lib.t2o_num_params(op, 8) for op in (-1, 0, 1, 2, 3, 5, 6, 7, 8, 9)] == [0, 1, 1, 1, 24, 8, 1, 1, 1, 3]
    assert lib.t2o_num_params(42, 8) == -1
    assert lib.t2o_workspace_bytes(16, 2048, 3072, 36) > 0 and lib.t2o_score_workspace_bytes(8, 100, 128, 128) > 0


def test_argument_validation_without_a_gpu():
    """Validation happens before any CUDA call, so the status codes are testable on a CPU box."""
    from t2onet_b200 import _lib
    lib = _lib.lib()
    ops, offs = _lib.int_array([0]), _lib.int_array([0])
    fake = ctypes.c_void_p(256)
    def fwd(n_ops=1, op=0, img=fake, pstride=1, L=8, flags=0, out=fake, mask=None, mask_ch=0):
        return lib.t2o_chain_forward(n_ops, _lib.int_array([op] * max(n_ops, 1)), _lib.int_array([0] * max(n_ops, 1)), img, mask,
                                     mask_ch, fake, pstride, None, out, None, 1, 8, 8, L, flags, None, 0, None)
    assert fwd(n_ops=0) == 1 and fwd(n_ops=9) == 1          # invalid chain length
    assert fwd(op=4) == 2                                    # inpaint unsupported
    assert fwd(op=42) == 1
    assert fwd(L=9) == 2                                     # curve_steps > 8
    assert fwd(op=3, pstride=3) == 1                         # parameter row too short
    assert fwd(img=None) == 1 and fwd(out=None) == 1
    assert fwd(mask=fake, mask_ch=2) == 1
    assert fwd(flags=2) == 1
    two = lib.t2o_chain_forward(2, _lib.int_array([6, 6]), _lib.int_array([0, 1]), fake, None, 0, fake, 2, None, fake, None,
                                1, 8, 8, 8, 0, None, 0, None)
    assert two == 2                                          # two stencils in one launch
    bwd = lib.t2o_chain_backward(1, ops, offs, fake, None, 0, fake, 1, None, None, None, fake, None, None, None,
                                 1, 8, 8, 8, None, 0, None)
    assert bwd == 1                                          # neither grad_out nor target+grad_l1
    bwd = lib.t2o_chain_backward(1, ops, offs, fake, None, 0, fake, 1, fake, None, None, fake, None, None, None,
                                 1, 8, 8, 8, None, 0, None)
    assert bwd == 3                                          # workspace missing
    # per-row chains: host-known rows are validated before any launch
    def rows_fwd(ops, K, pstride=48, slot=24, host=True, B=2):
        return lib.t2o_rows_forward(K, fake, _lib.int_array(ops) if host else None, slot, fake, None, 0, fake, pstride, None,
                                    fake, None, None, B, 8, 8, 8, None, 0, None)
    assert rows_fwd([0, 4, 1, 2], 2) == 2                    # inpaint in a row
    assert rows_fwd([0, 1, 6, 6], 2) == 2                    # two stencils in one row
    assert rows_fwd([0, 42, 1, 2], 2) == 1
    assert rows_fwd([0, 1, 1, 2], 2, pstride=24) == 1        # parameter table narrower than K slots
    assert rows_fwd([0, 1, 1, 2], 2, slot=8, pstride=16) == 2   # slot too small for the color curve
    assert rows_fwd([0, 1, 1, 2], 2, host=False) == 1        # device-only ids need K == 1
    rows_bwd = lib.t2o_rows_backward(2, fake, _lib.int_array([0, 0, 1, 2]), 24, fake, None, 0, fake, 48, fake, None, None,
                                     fake, None, None, None, None, 2, 8, 8, 8, fake, 1 << 30, None)
    assert rows_bwd == 2                                     # an operator type twice in one row of a backward call


def test_cpu_tensors_are_rejected_loudly():
    import t2onet_b200 as T
    with pytest.raises(T.T2OError):
        T.functional.chain(torch.rand(1, 3, 4, 4), [0], [torch.zeros(1, 1)])
    ex = T.Executor(T.default_options())
    with pytest.raises(T.T2OError):
        ex.execute(torch.rand(1, 3, 4, 4), 1, None, specified_param=torch.zeros(1, 1))
    out, p = ex.execute(torch.rand(2, 3, 4, 4), -1, None)    # identity never touches the library
    assert tuple(p.shape) == (2, 24)


# ------------------------------------------------------------------ host logic
def test_pack_params_and_segments():
    from t2onet_b200 import functional as TF
    B = 3
    ps = [torch.full((B, 1), 1.0), torch.full((B, 24), 2.0), torch.full((B, 8), 3.0), torch.full((B, 5), 4.0)]
    packed, offs, stride = TF.pack_params([0, 3, 5, 6], ps, B, 'cpu')
    assert offs == [0, 1, 25, 33] and stride == 34 and tuple(packed.shape) == (B, 34)
    assert packed[0, 0] == 1 and packed[0, 1] == 2 and packed[0, 25] == 3 and packed[0, 33] == 4
    assert TF.split_segments([0, 1, 2, 3, 5, 6]) == [[0, 1, 2, 3, 4, 5]]
    assert TF.split_segments([6, 1, 6, 5]) == [[0, 1], [2, 3]]
    assert TF.split_segments(list(range(10))) == [list(range(8)), [8, 9]]
    assert TF.split_segments([6, 6, 6]) == [[0], [1], [2]]
    assert TF.split_segments([0, 0, 1, 5, 1]) == [[0], [1, 2, 3], [4]]      # each operator type once per launch
    assert TF.split_segments([-1, 0, -1, 0]) == [[0, 1, 2], [3]]
    assert [TF.num_params(o) for o in (-1, 0, 3, 5, 9)] == [0, 1, 24, 8, 3]


def test_executor_module_layout_matches_reference():
    import t2onet_b200 as T
    torch.manual_seed(0)
    ex = T.Executor(T.default_options())
    keys = list(ex.state_dict().keys())
    assert keys[:4] == ['brightness_op.fc1.weight', 'brightness_op.fc1.bias', 'brightness_op.fc2.weight', 'brightness_op.fc2.bias']
    assert [k.split('.')[0] for k in keys[::4]] == ['brightness_op', 'sharpness_op', 'color_op', 'contrast_op', 'inpaint_op',
                                                    'white_op', 'saturation_op', 'tone_op']
    assert sum(p.numel() for p in ex.parameters()) == 2120742      # Executor 2.12 M parameters
    assert [op.num_op_param for op in ex.ops] == [1, 1, 1, 24, 1, 8, 1, 1]
    assert ex.ops[2].get_param_range() == (0.8, -0.2, 0) and ex.ops[5].get_param_range() == (2, 0.5, 1.25)
    f = torch.randn(5, 512)
    from oracle import ops as O
    for op_ind in (0, 1, 2, 3, 5, 6, 7):
        Op = ex.ops[op_ind]
        h = torch.nn.functional.leaky_relu(f @ Op.fc1.weight.t() + Op.fc1.bias)
        want = O.regress(op_ind, h @ Op.fc2.weight.t() + Op.fc2.bias)
        assert torch.allclose(Op.extract_parameters(f), want, atol=1e-6)


@pytest.mark.parametrize('n,x0', [(1, [0.0]), (1, [1.0]), (8, [1.0] * 8), (24, [1.0] * 24), (3, [0.0, 1.0, 0.0])])
def test_nelder_mead_coroutine_is_scipy_exact(n, x0):
    from scipy.optimize import minimize
    from t2onet_b200.nelder_mead import nelder_mead, run_lockstep
    rng = np.random.default_rng(n)
    A, c = rng.normal(size=(n, n)), rng.normal(size=n)
    f = lambda x: float(np.float32(np.abs(A @ x - c).sum() + 0.1 * np.sin(3 * x).sum()))   # noqa: E731
    trace = []
    res = minimize(lambda x: (trace.append(x.copy()), f(x))[1], np.array(x0), method='Nelder-Mead')
    mine = []

    def score(keys, pts):
        mine.append(pts[0].copy())
        return [f(p) for p in pts]
    r = run_lockstep({0: nelder_mead(np.array(x0))}, score)[0]
    assert len(trace) == len(mine) and all(np.array_equal(a, b) for a, b in zip(trace, mine))
    assert np.array_equal(res.x, r.x) and res.nfev == r.nfev and res.nit == r.nit and bool(res.success) == r.success


def test_lockstep_runs_many_fits_with_one_scoring_call_per_round():
    from t2onet_b200.nelder_mead import nelder_mead, run_lockstep
    calls = []
    targets = {i: np.full(d, 0.3 * (i + 1)) for i, d in enumerate((1, 1, 8, 1))}

    def score(keys, pts):
        calls.append(len(keys))
        return [float(((p - targets[k]) ** 2).sum()) for k, p in zip(keys, pts)]
    res = run_lockstep({i: nelder_mead(np.zeros(len(t))) for i, t in targets.items()}, score)
    assert calls[0] == 4 and calls[-1] == 1 and max(calls) == 4
    for i, t in targets.items():
        assert res[i].nfev <= 200 * len(t)
        if len(t) == 1:                      # (a zero start in 8-D gives a 2.5e-4 simplex: scipy does not converge either)
            assert np.abs(res[i].x - t).max() < 2e-3


# ------------------------------------------------------------------ the kernels' arithmetic, on the host
@pytest.fixture(scope='module')
def hostcheck():
    src = os.path.join(ROOT, 'tests', 'hostcheck', 'hostcheck.cpp')
    so = os.path.join(ROOT, 'tests', 'hostcheck', 'libhostcheck.so')
    hdr = os.path.join(ROOT, 't2onet_b200', 'csrc', 't2o_math.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-ffp-contract=fast', '-march=native', '-shared', '-fPIC', '-o', so, src], check=True)
    lib = ctypes.CDLL(so)

    def P(a, t=ctypes.c_float):
        return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))

    def run(ops, img, params, mask=None, grad_out=None, target=None, grad_l1=None, L=8):
        B, _, H, W = img.shape
        poff, off = [], 0
        for p in params:
            poff.append(off)
            off += p.shape[1]
        pst = max(off, 1)
        packed = np.zeros((B, pst), np.float32)
        for o, p in zip(poff, params):
            packed[:, o:o + p.shape[1]] = p
        out, l1 = np.zeros_like(img), np.zeros(B, np.float32)
        gp, gi = np.zeros((B, pst), np.float32), np.zeros_like(img)
        lib.hc_chain(len(ops), P(np.array(ops, np.int32), ctypes.c_int), P(np.array(poff, np.int32), ctypes.c_int), P(img),
                     P(mask), 0 if mask is None else mask.shape[1], P(packed), pst, P(grad_out), P(target), P(grad_l1),
                     P(out), P(l1) if target is not None else None, P(gp), P(gi), B, H, W, L)
        return out, l1, [gp[:, o:o + p.shape[1]] for o, p in zip(poff, params)], gi
    return run


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.mark.parametrize('op', [0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12])
def test_kernel_math_single_ops_vs_reference_golden(hostcheck, golden_dir, op):
    G = dict(np.load(os.path.join(golden_dir, 'single_ops.npz')))
    G.update(np.load(os.path.join(golden_dir, 'single_ops_ext.npz')))
    for variant in ('n_none', 'n_m1', 'n_m3', 'w_none'):
        key = 'op%d_%s' % (op, variant)
        mk = variant.split('_')[1]
        mask = None if mk == 'none' else np.ascontiguousarray(G['mask1' if mk == 'm1' else 'mask3'])
        out, l1, gps, gi = hostcheck([op], np.ascontiguousarray(G['img']), [G[key + '_param']], mask,
                                     grad_out=np.ascontiguousarray(G['wgt']), target=np.ascontiguousarray(G['target']))
        assert np.abs(out - G[key + '_out']).max() <= 1e-5
        assert np.allclose(l1, G[key + '_l1sum'], rtol=3e-6)
        if op != 7:
            assert _rel(gps[0], G[key + '_gparam']) <= 1e-4
        nimg = 3 if op not in (0, 2) else 2        # image 2 = two-channel ties, see DESIGN.md section 3
        assert _rel(gi[:nimg], G[key + '_gimg'][:nimg]) <= 1e-4


@pytest.mark.parametrize('name', ['c6', 'c6r', 'c3', 'c2'])
def test_kernel_math_chains_vs_reference_golden(hostcheck, golden_dir, name):
    C = np.load(os.path.join(golden_dir, 'chains.npz'))
    ops = [int(v) for v in C[name + '_ops']]
    img = np.ascontiguousarray(C['img'])
    params = [C['%s_param%d' % (name, k)] for k in range(len(ops))]
    gl1 = np.full(img.shape[0], 1.0 / img.size, np.float32)
    out, l1, gps, gi = hostcheck(ops, img, params, None, None, np.ascontiguousarray(C[name + '_target']), gl1)
    assert np.abs(out - C[name + '_out']).max() <= 1e-5
    assert abs(l1.sum() / img.size - float(C[name + '_l1mean'])) <= 1e-5
    for k in range(len(ops)):
        assert _rel(gps[k], C['%s_gparam%d' % (name, k)]) <= 1e-4
    assert _rel(gi, C[name + '_gimg']) <= 1e-4


# ------------------------------------------------------------------ multi-rank plumbing on gloo (world size 2)
def _dist_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from t2onet_b200 import dist as D
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    try:
        # image-sharded planning: every rank ends with the full ordered list
        pairs = [(i, i * 10) for i in range(7)]
        recs = D.plan_dataset(pairs, None, lambda a, b, ex: {'item': a, 'val': a + b, 'rank': dist.get_rank()})
        assert [r['item'] for r in recs] == list(range(7))
        assert [r['rank'] for r in recs] == [i % world for i in range(7)]
        # the same in lock-step batches (planner.beam_search_batch wants many pairs in flight per rank)
        calls = []

        def batch_fn(idx, items, ex):
            calls.append(list(idx))
            return [{'item': a, 'val': a + b, 'rank': dist.get_rank()} for a, b in items]
        recs = D.plan_dataset_batched(pairs, None, batch_fn, batch=3)
        assert [r['item'] for r in recs] == list(range(7)) and [r['val'] for r in recs] == [11 * i for i in range(7)]
        assert [r['rank'] for r in recs] == [i % world for i in range(7)]
        assert all(len(c) <= 3 for c in calls) and sorted(sum(calls, [])) == list(range(rank, 7, world))
        # candidate-sharded selection: packed-key all_reduce(MIN)
        g = torch.Generator().manual_seed(5)
        scores = torch.rand(4, 10, generator=g)                 # the same global table on both ranks
        ids = torch.arange(10).repeat(4, 1)
        mine = slice(rank * 5, rank * 5 + 5)
        best, bid = D.best_candidate(scores[:, mine], ids[:, mine])
        assert torch.equal(best, scores.min(1).values) and torch.equal(bid, scores.argmin(1))
        # ties on the score resolve to the smaller candidate id on every rank
        tied = torch.full((1, 5), 0.25)
        best, bid = D.best_candidate(tied, torch.arange(5).view(1, 5) + 5 * rank)
        assert bid.item() == 0 and best.item() == 0.25
        # the candidate-sharded planner's per-step agreement (planner._step_minima_sharded): every rank contributes the
        # candidates it fitted (problems rank, rank + R, ...), every rank ends with each live pair's minimum and its problem
        from t2onet_b200 import planner as P
        problems = [(0, 0, 3, 0, None), (0, 1, 3, 0, None), (1, 0, 5, 0, None), (1, 2, 5, 0, None), (2, 6, 5, 1, None)]
        dists = [0.5, 0.25, 0.75, 0.125, 0.0625]
        res = P._step_minima_sharded(dists, problems, [3, 5], torch.device('cpu'), None)
        assert res == {3: (0.25, 1), 5: (0.0625, 4)}, res
        q.put((rank, 'ok'))
    except Exception as exc:            # pragma: no cover
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_distributed_plumbing_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, 'ok'), (1, 'ok')], res


def test_score_key_packing_orders_like_floats():
    from t2onet_b200 import dist as D
    s = torch.tensor([0.0, 1e-30, 0.5, 0.50000006, 3.0, 1e30])
    keys = D.pack_score_keys(s, torch.zeros(6, dtype=torch.int64))
    assert torch.equal(torch.argsort(keys), torch.arange(6))
    sc, ids = D.unpack_score_keys(D.pack_score_keys(s, torch.arange(6)))
    assert torch.equal(sc, s) and torch.equal(ids, torch.arange(6))
    assert D.shard_indices(7, 1, 3) == [1, 4] and D.shard_indices(2, 3, 4) == []


# ------------------------------------------------------------------ planner records on disk (t2onet_b200/plans.py)
def test_plan_records_encode_like_the_reference_reader(tmp_path):
    """encode_plan / read_plan == FiveKAct.get_act of the unmodified reference (datasets/FiveKdataset.py:86-113) on the
    records of tests/golden/plans.json (oracle/make_plans_golden.py): operator ids, truncation, parameter normalisation."""
    import json
    from t2onet_b200 import plans
    cases = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'plans.json')))
    assert len({c['trunc_len'] for c in cases}) >= 3
    for i, c in enumerate(cases):
        op_seq, params, trunc, seq = plans.encode_plan(json.loads(json.dumps(c['record'])))
        assert [int(v) for v in op_seq] == c['op_seq'] and trunc == c['trunc_len'] and len(seq) == trunc
        assert np.array_equal(params, np.array(c['params'], dtype=np.float32))
        # the writer's layout is the reader's: <save_dir>/<phase><i>/<i:05d>.json
        rec = c['record']
        plans.write_plan(str(tmp_path), 'train', i, rec['request'], None, None, rec['operation sequence'], [],
                         rec['init distance'], write_images=False)
        op_seq2, params2, trunc2, _ = plans.read_plan(str(tmp_path), 'train', i)
        assert np.array_equal(op_seq2, op_seq) and np.array_equal(params2, params) and trunc2 == trunc
        with open(os.path.join(str(tmp_path), 'train%d' % i, '%05d.json' % i)) as f:
            assert json.load(f) == rec


def test_analyze_traj_rule():
    from t2onet_b200 import plans
    assert plans.analyze_traj([1.0, 0.5, 0.2, 0.199, 0.1]) == 2       # third step improves by < 1 % of the initial distance
    assert plans.analyze_traj([1.0, 0.999]) == 1                       # never 0
    assert plans.analyze_traj([1.0, 0.5, 0.2]) == 2                    # every step counts
    img = torch.rand(1, 3, 5, 7)
    bgr = plans.tensor2img(img)
    assert bgr.shape == (5, 7, 3) and bgr.dtype == np.uint8
    assert np.array_equal(bgr[:, :, ::-1], (img[0].permute(1, 2, 0) * 255).numpy().astype(np.uint8))
    assert torch.equal(plans.img2tensor(bgr)[0], torch.from_numpy(bgr[:, :, ::-1].transpose(2, 0, 1).copy()) / 255)
