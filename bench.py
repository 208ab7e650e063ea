#!/usr/bin/env python
"""bench.py -- the headline benchmark of the T2ONet hot path on B200 (contract: see DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2|c1|c3]

Metric (BASELINE.json): edited Mpixel/s of the operator chain forward+backward (& planner candidates/s).  One "step" =
one pass of the fused chain  [brightness, contrast, saturation, color, tone, sharpness]  over one synthetic batch:
edited image + per-image L1 to the target + gradients to all 36 operator parameters, ONE kernel launch.
Default workload = BASELINE config 4 (batch 16 of 3x2048x3072: the roofline configuration, 3.6 GB >> L2); the line also
carries config 2 (batch 64 of 3x128x128, the seq2seqL1 training shape) under "c2" and the planner (config 3's shape)
under "planner", each with its own CPU baseline.

  value  : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    : the same step through the public API from pinned HOST buffers: 8-bit images host -> device (the form the
           reference's images have on the host, utils/visual_utils.py:61-70), x / 255 on the device, fused step,
           device -> host of the L1 terms and parameter gradients
  roofline / cpu_baseline / clocks : see DESIGN.md
`--impl reference` times the reference's own CPU implementation on the host cores for the same config: the unmodified
reference operators byte-compiled into oracle/_ref (kind "reference") when present, else the oracle port (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CHAIN = [0, 1, 2, 3, 5, 6]
CHAIN_NAMES = ['brightness', 'contrast', 'saturation', 'color', 'tone', 'sharpness']
WORKLOADS = {
    'c1': dict(B=1, H=512, W=512, desc='C1: single 3x512x512 image, 6-op chain fwd+L1+bwd'),
    'c2': dict(B=64, H=128, W=128, desc='C2: batch 64 of 3x128x128 (seq2seqL1 training shape), 6-op chain fwd+L1+bwd to all operator params'),
    'c4': dict(B=16, H=2048, W=3072, desc='C4: batch 16 of 3x2048x3072, 6-op chain fwd+L1+bwd (bandwidth stress)'),
}
# bounded CPU sample of each workload (one reference-arm step / the cpu_baseline sample): the full C2 batch, one image
# of C1, a quarter image of C4 (autograd keeps ~50 planes per operator alive: a full 3x2048x3072 image needs > 10 GB)
CPU_SAMPLE = {'c1': (1, 512, 512), 'c2': (64, 128, 128), 'c4': (1, 1024, 1536)}
L2_BYTES = 126 * 1024 * 1024
BYTES_PER_PX_FUSED = 36       # read img 12 + read target 12 + write out 12 (one fused fwd+bwd launch), DESIGN.md section 4


def make_params(B, gen, device):
    """Parameter distributions of SURVEY.md section 8(d)."""
    u = lambda n: torch.rand(B, n, generator=gen)   # noqa: E731
    ps = [u(1) * 0.6 - 0.3, u(1) - 0.5, u(1) - 0.2, 0.9 + 0.2 * u(24), 0.5 + 1.5 * u(8), u(1) * 1.5]
    return [p.to(device) for p in ps]


def make_batch(B, H, W, seed, device):
    gen = torch.Generator().manual_seed(seed)
    if device == 'cpu' or B * H * W <= 4 * 1024 * 1024:
        img = torch.rand(B, 3, H, W, generator=gen).to(device)
        tgt = torch.rand(B, 3, H, W, generator=gen).to(device)
    else:
        dgen = torch.Generator(device=device).manual_seed(seed)
        img = torch.rand(B, 3, H, W, generator=dgen, device=device)
        tgt = torch.rand(B, 3, H, W, generator=dgen, device=device)
    return img, tgt, make_params(B, gen, device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                clocks.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        clocks.sort()
        med = clocks[len(clocks) // 2] if clocks else None
        return {'sm_mhz': med, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(clocks)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------ reference arm
def workload_config(wl):
    """The `config` object both arms print (identical keys and values)."""
    return {'workload': wl['desc'], 'chain': CHAIN_NAMES, 'batch_per_gpu': wl['B'], 'H': wl['H'], 'W': wl['W']}


_REF = {}


def reference_backend():
    """-> (kind, step(img, tgt, params) -> loss).  kind "reference": the reference's own Executor / operators
    (models/operators.py, executors/executor.py, byte-compiled into oracle/_ref by oracle/build_ref.py; kornia = the
    oracle's restatement) with torch autograd; kind "port": the oracle's restatement of them (oracle/ops.py)."""
    if 'step' in _REF:
        return _REF['kind'], _REF['step']
    try:
        from oracle import ref_shims
        if not ref_shims.available():
            raise RuntimeError('no reference tree')
        R = ref_shims.load()
        torch.manual_seed(10)
        ex = R.executor.Executor(R.options())

        def step(img, tgt, params):
            ps = [p.clone().requires_grad_() for p in params]
            x = img
            for op, p in zip(CHAIN, ps):
                x = ex.execute(x, op, None, specified_param=p)[0]
            loss = (x - tgt).abs().mean()                         # experiments/t2onet/train_seq2seqL1.py:85
            loss.backward()
            return loss.item()
        _REF.update(kind='reference', step=step)
    except Exception as exc:
        print('[bench] reference tree unavailable (%r): timing the oracle port' % (exc,), file=sys.stderr)
        from oracle import ops as O

        def step(img, tgt, params):
            ps = [p.clone().requires_grad_() for p in params]
            out = O.chain(img, CHAIN, ps)
            loss = (out - tgt).abs().mean()
            loss.backward()
            return loss.item()
        _REF.update(kind='port', step=step)
    return _REF['kind'], _REF['step']


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    kind, step = reference_backend()
    B, H, W = wl['B'], wl['H'], wl['W']
    sB, sH, sW = CPU_SAMPLE[args.workload]
    img, tgt, params = make_batch(sB, sH, sW, 10 + 2000, 'cpu')
    for _ in range(max(1, args.warmup)):
        step(img, tgt, params)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(img, tgt, params)
    dt = time.perf_counter() - t0
    mpix = sB * sH * sW * args.steps / dt / 1e6
    sample = '%d timed steps (+ %d warm-up) of %dx3x%dx%d per step (the configured batch is %dx3x%dx%d); %s' % (
        args.steps, max(1, args.warmup), sB, sH, sW, B, H, W,
        'unmodified reference operators (oracle/_ref) with torch CPU autograd' if kind == 'reference'
        else 'oracle port of the reference operators (oracle/ops.py), torch CPU autograd')
    line = {
        'impl': 'reference', 'metric': 'edited Mpixel/s (op-chain fwd+bwd)', 'value': mpix, 'unit': 'Mpixel/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl),
        'cpu_baseline': {'value': mpix, 'unit': 'Mpixel/s', 'cores': torch.get_num_threads(), 'kind': kind, 'sample': sample},
        'e2e': {'value': mpix, 'unit': 'Mpixel/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def timed_region(fn, steps, stream_sync):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_sync()
    start.record()
    for i in range(steps):
        fn(i)
    end.record()
    stream_sync()
    return start.elapsed_time(end)


def measure_fused(TF, dev, wl, steps, warmup, rank, use_graph, barrier):
    """The fused step (forward + L1 + backward, one launch) over rotating device-resident batches.
    -> dict(ms, ms_eager, nb, graph, batches, packed, batch_bytes)"""
    B, H, W = wl['B'], wl['H'], wl['W']
    px_step = B * H * W
    batch_bytes = 3 * px_step * 12
    nb = max(2, min(16, (2 * L2_BYTES + batch_bytes - 1) // batch_bytes + 1)) if batch_bytes < 2 * L2_BYTES else 1
    batches = [make_batch(B, H, W, 10 + 2000 + 17 * i + 1000 * rank, dev) for i in range(nb)]
    # one prepared step per rotating batch: its own parameter table and its own output buffers, so inputs AND
    # outputs rotate through more memory than the L2 holds
    packed = [torch.cat(b[2], 1).contiguous() for b in batches]
    fused = [TF.FusedStep(CHAIN, B, H, W, dev, want_out=True, want_grad_img=False, reuse_outputs=True) for _ in range(nb)]

    def step(i):
        j = i % nb
        return fused[j](batches[j][0], packed[j], batches[j][1])

    for i in range(max(3, warmup, nb)):
        step(i)
    # the timed region replays a CUDA graph of the step launches (one kernel node per step) unless --no-graph
    graph = None
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(nb):
                    step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(steps):
                    step(i)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:      # report, never hide
            print('[bench] CUDA graph capture failed, timing eager launches: %r' % (exc,), file=sys.stderr)
            graph = None
    return dict(step=step, graph=graph, nb=nb, batches=batches, packed=packed, batch_bytes=batch_bytes, px_step=px_step)


def measure_e2e(TF, dev, wl, batches, packed, steps, barrier, u8):
    """The same step through the public API from pinned HOST memory.  Every step copies its inputs host -> device
    (u8: the 8-bit images + visual_utils.u8_to_float on the device; else the float32 tensors), runs the prepared step and
    copies the loss terms and parameter gradients back; two steps are in flight (two streams with their own device
    buffers), the host consumes the result of step i before it issues step i + 2.  -> (ms, steps, h2d bytes, d2h bytes)"""
    from t2onet_b200 import visual_utils as V
    B, H, W = wl['B'], wl['H'], wl['W']
    px_step = B * H * W
    NSLOT = int(os.environ.get('T2O_E2E_SLOTS', 2))
    nh = min(2, len(batches))
    if u8:
        # image and target of a step sit back to back in ONE pinned buffer: one host -> device copy per step
        hboth = [torch.stack([V.float_to_u8(b[0]).cpu(), V.float_to_u8(b[1]).cpu()]).pin_memory() for b in batches[:nh]]
        stage = [torch.empty(2, B, 3, H, W, dtype=torch.uint8, device=dev) for _ in range(NSLOT)]
    else:
        himg = [b[0].cpu().pin_memory() for b in batches[:nh]]
        htgt = [b[1].cpu().pin_memory() for b in batches[:nh]]
    hpar = [p.cpu().pin_memory() for p in packed[:nh]]
    streams = [torch.cuda.Stream() for _ in range(NSLOT)]
    dimg = [torch.empty_like(batches[0][0]) for _ in range(NSLOT)]
    dtgt = [torch.empty_like(batches[0][1]) for _ in range(NSLOT)]
    dpar = [torch.empty_like(packed[0]) for _ in range(NSLOT)]
    res_l1 = [torch.empty(B, pin_memory=True) for _ in range(NSLOT)]
    res_gp = [torch.empty(B, 36, pin_memory=True) for _ in range(NSLOT)]
    fused_e2e = [TF.FusedStep(CHAIN, B, H, W, dev, want_out=True, want_grad_img=False, reuse_outputs=True) for _ in range(NSLOT)]
    done = [None] * NSLOT
    losses = []

    def e2e_step(i):
        s, j = i % NSLOT, i % nh
        if done[s] is not None:
            done[s].synchronize()                       # the caller consumes the loss of step i - NSLOT
            losses.append(float(res_l1[s][0]))
        with torch.cuda.stream(streams[s]):
            if u8:
                stage[s].copy_(hboth[j], non_blocking=True)
                V.u8_to_float(stage[s][0], out=dimg[s])
                V.u8_to_float(stage[s][1], out=dtgt[s])
            else:
                dimg[s].copy_(himg[j], non_blocking=True)
                dtgt[s].copy_(htgt[j], non_blocking=True)
            dpar[s].copy_(hpar[j], non_blocking=True)
            _, l1, gp, _ = fused_e2e[s](dimg[s], dpar[s], dtgt[s])
            res_l1[s].copy_(l1, non_blocking=True)
            res_gp[s].copy_(gp, non_blocking=True)
            done[s] = torch.cuda.Event()
            done[s].record()

    def e2e_region(n):
        cur = torch.cuda.current_stream()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for st in streams:
            st.wait_event(start)
        for i in range(n):
            e2e_step(i)
        for st in streams:
            cur.wait_stream(st)
        end.record()
        barrier()
        for s in range(NSLOT):
            done[s] = None
        return start.elapsed_time(end)

    e2e_region(4)
    ms = e2e_region(steps)
    h2d = 2 * px_step * (3 if u8 else 12) + hpar[0].numel() * 4
    d2h = B * 4 + B * 36 * 4
    return ms, h2d, d2h, (3 if u8 else 1)      # launches per step: two conversions + the fused step


def kernel_counters(key):
    """Counters of the dominant kernel from the committed ncu capture of this workload (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)[key]
    except Exception:
        return {}


def roofline_block(workload, px_step, kernel_ms, clocks):
    peak, peak_src = measured_peak_hbm()
    achieved = BYTES_PER_PX_FUSED * px_step / (kernel_ms * 1e-3) / 1e9
    kc = kernel_counters(workload)
    traffic = kc['dram_read_bytes'] + kc['dram_write_bytes'] if 'dram_read_bytes' in kc else None
    blk = {'bound': 'hbm', 'kernel': kc.get('kernel', 'step_sharp_kernel<4,0,256,0,SP_C6>') + ' (fused forward + L1 + backward, one launch per step)',
           'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
           'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram read + write)', 'traffic_source': kc.get('source'),
           'algorithmic_bytes': BYTES_PER_PX_FUSED * px_step, 'peak_source': peak_src,
           'algorithmic_bytes_per_px': BYTES_PER_PX_FUSED, 'frac_of_nominal_8TBs': achieved / 8000.0}
    # the second roofline: the kernel is FP32-issue-bound -- thread-instructions per pixel from the ncu
    # capture against 148 SMs x 4 schedulers x 32 lanes per clock at the SM clock sampled during the timed region
    ipp = kc.get('thread_instr_per_px')
    mhz = (clocks or {}).get('sm_mhz') or (clocks or {}).get('sm_max_mhz')
    if ipp and mhz:
        issue_peak = 148 * 4 * 32 * mhz * 1e6
        achieved_issue = ipp * px_step / (kernel_ms * 1e-3)
        blk['issue'] = {'thread_instr_per_px': ipp, 'achieved_thread_instr_per_s': achieved_issue, 'peak_thread_instr_per_s': issue_peak,
                        'frac': achieved_issue / issue_peak, 'sm_mhz': mhz, 'source': kc.get('source'),
                        'note': 'issue-slot roofline: the time the same instruction stream needs at one warp-instruction per '
                                'scheduler and clock is frac x the measured time'}
    return blk


def run_ours(args, wl):
    import t2onet_b200.functional as TF
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def run_workload(key, w, steps, warmup, sample_clocks):
        m = measure_fused(TF, dev, w, steps, warmup, rank, not args.no_graph, barrier)
        sampler = ClockSampler(local)
        if sample_clocks and rank == 0:
            sampler.start()
        if m['graph'] is not None:
            ms = timed_region(lambda i: m['graph'].replay(), 1, barrier)      # one replay = exactly `steps` step launches
        else:
            ms = timed_region(m['step'], steps, barrier)
        clocks = sampler.stop() if sample_clocks and rank == 0 else None
        ms_eager = timed_region(m['step'], steps, barrier)
        e2e_steps = max(4, min(steps, 50 if m['px_step'] < (1 << 24) else 10))
        ms_e2e, h2d, d2h, per_step_launches = measure_e2e(TF, dev, w, m['batches'], m['packed'], e2e_steps, barrier, u8=True)
        ms, ms_eager, ms_e2e = max_over_ranks(ms, ms_eager, ms_e2e)
        res = dict(m, ms=ms, ms_eager=ms_eager, clocks=clocks, steps=steps, e2e_ms=ms_e2e, e2e_steps=e2e_steps, h2d=h2d, d2h=d2h,
                   value=world * m['px_step'] * steps / (ms * 1e-3) / 1e6,
                   e2e_value=world * m['px_step'] * e2e_steps / (ms_e2e * 1e-3) / 1e6, e2e_launches_per_step=per_step_launches)
        return res

    main = run_workload(args.workload, wl, args.steps, args.warmup, True)
    second = None
    if args.workload != 'c2' and not args.no_extras:
        # free the main workload's buffers first (C4 holds ~5 GB per rotating batch)
        for k in ('batches', 'packed', 'step', 'graph'):
            main[k] = None
        torch.cuda.empty_cache()
        second = run_workload('c2', WORKLOADS['c2'], max(args.steps, 50), args.warmup, False)

    planner_line = None
    if world > 1 and not args.no_extras:
        # planner throughput scaling: every rank plans its own 64 pairs (image sharding, no data-path communication)
        barrier()
        err = None
        try:
            sec, cand, steps = planner_e2e_run(dev, 3010 + 17 * rank)
        except Exception as exc:   # report, never hide -- and still take part in the collectives below
            sec, cand, steps, err = 0.0, 0, 0.0, repr(exc)
        t = torch.tensor([sec, float(cand), steps, 0.0 if err is None else 1.0], device=dev, dtype=torch.float64)
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if tsum[3].item() > 0:
            planner_line = {'error': err or 'a rank failed'}
        else:
            planner_line = planner_e2e_line(tmax[0].item(), int(tsum[1].item()), tsum[2].item() / world, n_gpus=world)
    shard_line = None
    if world > 1 and not args.no_extras:
        # candidate-sharded planning: every rank holds the SAME 8 pairs, the fits of a step are split
        # over the ranks, the fit table is all-gathered and the step's best candidate per pair agreed by all_reduce(MIN) over NCCL
        barrier()
        try:
            import t2onet_b200 as T
            from t2onet_b200 import planner
            names = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
            exe = T.Executor(T.default_options()).to(dev)
            img, tgt, _ = make_batch(8, 128, 128, 3010, dev)
            best = None
            for rep in range(3):
                cnt = [0]
                barrier()
                t0 = time.time()
                res = planner.beam_search_batch(img, tgt, exe, 8, CHAIN, names, 6, 1e-2, counter=cnt, shard_fits=True)
                torch.cuda.synchronize()
                dt = time.time() - t0
                if rep > 0 and (best is None or dt < best):
                    best = dt
            sig = torch.tensor([float(sum(a[2] for r in res for seq in r[0] for a in seq))], device=dev, dtype=torch.float64)
            sigs = [torch.zeros_like(sig) for _ in range(world)]
            dist.all_gather(sigs, sig)
            t = torch.tensor([best], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            shard_line = {'workload': '8 pairs of 3x128x128 (the same on every rank), beam 8, ops [0,1,2,3,5,6]: the fits of every step split over the '
                                      'ranks, one all_gather of a 208-byte record per fit and one all_reduce(MIN) of 8 bytes per pair and step',
                          'n_gpus': world, 'seconds': t.item(), 'pairs_per_s': 8 / t.item(), 'candidates': cnt[0], 'candidates_per_s': cnt[0] / t.item(),
                          'identical_on_every_rank': bool(all(torch.equal(sigs[0], v) for v in sigs))}
        except Exception as exc:
            shard_line = {'error': repr(exc)}
    gier_line = None
    if world > 1 and not args.no_extras:
        barrier()
        err = None
        try:
            sec, cand, steps = planner_gier_run(dev, 5010 + 17 * rank)
        except Exception as exc:
            sec, cand, steps, err = 0.0, 0, 0.0, repr(exc)
        t = torch.tensor([sec, float(cand), steps, 0.0 if err is None else 1.0], device=dev, dtype=torch.float64)
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        gier_line = {'error': err or 'a rank failed'} if tsum[3].item() > 0 else \
            planner_gier_line(tmax[0].item(), int(tsum[1].item()), tsum[2].item() / world, n_gpus=world)
    ddp_line = None
    if world > 1 and not args.no_extras:
        try:
            ddp_line = ddp_actor_leg(dev, dist, rank, world)
        except Exception as exc:
            ddp_line = {'error': repr(exc)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    kernel_ms = main['ms'] / args.steps
    nb, batch_bytes = main['nb'], main['batch_bytes']
    line = {
        'metric': 'edited Mpixel/s (op-chain fwd+bwd)', 'value': main['value'], 'unit': 'Mpixel/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': kernel_ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl),
        'run': {'l2_policy': ('rotating %d distinct batches (%.0f MB) > 126 MB L2' % (nb, nb * batch_bytes / 1e6)) if nb > 1
                else 'one batch of %.0f MB >> 126 MB L2' % (batch_bytes / 1e6),
                'sharding': 'images sharded across ranks, no data-path collective',
                'launch': 'CUDA graph replay of %d step launches' % args.steps if not args.no_graph else 'eager launches',
                'eager_ms_per_step': main['ms_eager'] / args.steps},
        'clocks': main['clocks'],
        'e2e': {'value': main['e2e_value'], 'unit': 'Mpixel/s', 'h2d_bytes_per_step': main['h2d'], 'd2h_bytes_per_step': main['d2h'],
                'ms_per_step': main['e2e_ms'] / main['e2e_steps'], 'steps': main['e2e_steps'],
                'how': '8-bit images + float32 parameters from pinned host buffers, x / 255 on the device (visual_utils.u8_to_float), '
                       'FusedStep, L1 terms + parameter gradients back to pinned host buffers; 2 steps in flight on 2 streams'},
        'gpu_launches': args.steps,
        'roofline': roofline_block(args.workload, main['px_step'], kernel_ms, main['clocks']),
    }
    if second is not None:
        c2ms = second['ms'] / second['steps']
        line['c2'] = {'config': workload_config(WORKLOADS['c2']), 'value': second['value'], 'unit': 'Mpixel/s', 'ms_per_step': c2ms,
                      'steps': second['steps'], 'eager_ms_per_step': second['ms_eager'] / second['steps'],
                      'l2_policy': 'rotating %d distinct batches (%.0f MB) > 126 MB L2' % (second['nb'], second['nb'] * second['batch_bytes'] / 1e6),
                      'e2e': {'value': second['e2e_value'], 'unit': 'Mpixel/s', 'h2d_bytes_per_step': second['h2d'],
                              'd2h_bytes_per_step': second['d2h'], 'ms_per_step': second['e2e_ms'] / second['e2e_steps']},
                      'roofline': roofline_block('c2', second['px_step'], c2ms, main['clocks'])}
    if world == 1 and not args.no_extras:
        line['cpu_baseline'] = cpu_baseline(args.workload, wl)
        if second is not None:
            line['c2']['cpu_baseline'] = cpu_baseline('c2', WORKLOADS['c2'], seconds=6.0)
        ex = extras(TF, dev, wl)
        line['planner'] = planner_block(ex)
        line['extras'] = ex
    if planner_line is not None:
        line['planner'] = {'metric': 'planner candidates/s', 'value': planner_line.get('candidates_per_s'), 'unit': 'candidates/s',
                           'e2e': planner_line}
    if shard_line is not None:
        line['planner_candidate_sharded'] = shard_line
    if gier_line is not None:
        line['planner_gier'] = gier_line
    if ddp_line is not None:
        line['ddp_actor'] = ddp_line
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def planner_block(ex):
    """The planner half of the metric (candidates/s) as a top-level object: the scorer's sweep, the public-API run and
    the reference's CPU planner beside it."""
    blk = {'metric': 'planner candidates/s', 'unit': 'candidates/s'}
    e2e, sweep = ex.get('planner_e2e', {}), ex.get('planner_scoring', {})
    blk['value'] = e2e.get('candidates_per_s')
    blk['e2e'] = e2e
    blk['scoring_sweep'] = sweep
    blk['gier'] = ex.get('planner_gier', {})
    try:
        cpu = planner_cpu_baseline()
        if e2e.get('candidates') and e2e.get('pairs_per_s') and cpu.get('value'):
            per_pair = e2e['candidates'] / float(PLANNER_M)
            cpu['pairs_per_s_extrapolated'] = cpu['value'] / per_pair
            cpu['extrapolation'] = 'candidates/s divided by the GPU run\'s %.0f candidates per pair' % per_pair
        blk['cpu_baseline'] = cpu
    except Exception as exc:
        blk['cpu_baseline'] = {'error': repr(exc)}
    return blk


def _planner_cpu_worker(job):
    """One process, one torch thread: the reference's beam_search (or the oracle port) on one synthetic pair, first step only."""
    seed, max_step = job
    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    names = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
    img, tgt, _ = make_batch(1, 128, 128, seed, 'cpu')
    cnt = [0]
    kind = 'port'
    t0 = time.perf_counter()
    try:
        from oracle import ref_shims
        if not ref_shims.available():
            raise RuntimeError('no reference tree')
        R = ref_shims.load()
        torch.manual_seed(10)
        ex = R.executor.Executor(R.options())
        mod = R.beam_search
        orig = mod.get_dist

        def counted(x1, x2, dist_type):
            cnt[0] += 1
            return orig(x1, x2, dist_type)
        mod.get_dist = counted
        t0 = time.perf_counter()
        try:
            mod.beam_search(img, tgt, None, ex, None, 8, CHAIN, names, max_step, 1e-2, 'L1', 'Nelder-Mead', replace=False)
        finally:
            mod.get_dist = orig
        kind = 'reference'
    except Exception:
        from oracle import ops as O
        from oracle import planner as OP
        cnt[0] = 0
        t0 = time.perf_counter()
        OP.beam_search(img, tgt, None, O.OracleExecutor(), None, 8, CHAIN, names, max_step, 1e-2, 'L1', 'Nelder-Mead', counter=cnt)
    return cnt[0], time.perf_counter() - t0, kind


def planner_cpu_baseline(max_procs=16):
    """The reference's own planner on the host cores: utils/beam_search.beam_search, beam 8, the six
    operators, Nelder-Mead, on synthetic 3x128x128 pairs of the GPU workload's kind -- one pair per process (the search of a
    pair is sequential: scipy's Nelder-Mead evaluates one candidate at a time), max_step = 1 (6 fits, ~7 000 candidates:
    the bounded sample), all processes timed together."""
    import multiprocessing as mp
    procs = max(1, min(max_procs, os.cpu_count() or 1))
    jobs = [(3010 + 7 * i, 1) for i in range(procs)]
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(procs) as pool:
        res = pool.map(_planner_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    cand = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    kind = res[0][2]
    return {'value': cand / busy, 'unit': 'candidates/s', 'cores': procs, 'kind': kind,
            'sample': '%d pairs of 3x128x128 (one per process, 1 torch thread each), %s beam_search beam 8, ops [0,1,2,3,5,6], '
                      'Nelder-Mead, first step only (max_step 1): %d candidates in %.1f s (slowest process; %.1f s wall with '
                      'process start-up)' % (procs, 'the unmodified reference\'s' if kind == 'reference' else 'the oracle port\'s',
                                             cand, busy, wall),
            'per_process_candidates_per_s': cand / sum(r[1] for r in res)}


def ddp_actor_leg(dev, dist, rank, world, B=64, H=128, W=128, iters=6):
    """BASELINE config 5 / SURVEY.md section 8e row 1: the seq2seqL1 training step (episode_forward + mean L1 + backward,
    experiments/t2onet/train_seq2seqL1.py:75-88) of the reference's own Actor on the new Executor, wrapped in
    DistributedDataParallel: batch 64 per GPU, the 22.17 M-parameter gradient all-reduce over NCCL.  Needs the compiled
    reference tree (oracle/_ref); -> per-step time, max over ranks."""
    from oracle import ref_shims
    if not ref_shims.available():
        return {'unavailable': 'oracle/_ref (the byte-compiled reference Actor) is not present'}
    import t2onet_b200 as T
    from torch.nn.parallel import DistributedDataParallel as DDP
    opt = ref_shims.actor_options()
    actor = ref_shims.build_actor(opt, T.Executor, seed=10).to(dev)
    ddp = DDP(_EpisodeL1(actor, opt), device_ids=[torch.device(dev).index], find_unused_parameters=True)
    g = torch.Generator().manual_seed(10 + 9000 + rank)
    x = torch.randint(4, 200, (B, opt.encoder_max_len), generator=g)
    x[:, 0], x[:, -1] = opt.start_id, opt.end_id
    img = torch.rand(B, 3, H, W, generator=g).to(dev)
    tgt = torch.rand(B, 3, H, W, generator=g).to(dev)
    x = x.to(dev)
    optim = torch.optim.Adam(ddp.parameters(), lr=1e-4)
    times = []
    for it in range(iters):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        optim.zero_grad(set_to_none=True)
        loss = ddp(x, img, tgt)
        loss.backward()
        optim.step()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    # rank-identical gradients after the all-reduce
    gsum = torch.stack([p.grad.double().sum() for p in ddp.parameters() if p.grad is not None]).sum().view(1)
    gall = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(gall, gsum)
    t = torch.tensor([min(times[2:])], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nparam = sum(p.numel() for p in ddp.parameters())
    return {'workload': 'reference Actor (models/actor.py, unmodified) on t2onet_b200.Executor under DistributedDataParallel: '
                        'episode_forward + mean L1 + backward + Adam, batch %d x 3x%dx%d per GPU' % (B, H, W),
            'n_gpus': world, 'ms_per_step': t.item(), 'images_per_s': world * B / t.item() * 1e3,
            'allreduce_bytes_per_step': nparam * 4, 'parameters': nparam,
            'gradients_identical_across_ranks': bool(all(torch.equal(gall[0], v) for v in gall)),
            'loss': float(loss.item())}


class _EpisodeL1(torch.nn.Module):
    """experiments/t2onet/train_seq2seqL1.py:75-85 as a module, so that DDP's forward hook sees the whole step."""

    def __init__(self, actor, opt):
        super().__init__()
        self.actor, self.opt = actor, opt

    def forward(self, x, img_x, img_y):
        _, pred_imgs, pred_ops, _ = self.actor.episode_forward(x, img_x, None)
        bs, max_len = pred_imgs.shape[:2]
        end = []
        for b in range(bs):
            idxs = (pred_ops[b] == self.opt.end_id).nonzero()
            end.append(pred_imgs[b, idxs[0][0] if len(idxs) > 0 else max_len - 1])
        return torch.abs(torch.stack(end) - img_y).mean()


def cpu_baseline(key, wl, seconds=12.0):
    """The reference's CPU operators (unmodified, oracle/_ref) -- or their oracle port -- with torch autograd on the host
    cores, bounded sample; all cores and one thread."""
    kind, step = reference_backend()
    B, H, W = CPU_SAMPLE[key]
    img, tgt, params = make_batch(B, H, W, 10 + 2000, 'cpu')

    def run(threads, budget):
        torch.set_num_threads(threads)
        step(img, tgt, params)
        t0 = time.perf_counter()
        n = 0
        while (time.perf_counter() - t0 < budget and n < 40) or n < 1:
            step(img, tgt, params)
            n += 1
        return B * H * W * n / (time.perf_counter() - t0) / 1e6, n
    cores = os.cpu_count() or 1
    v_all, n_all = run(cores, seconds)
    v_one, n_one = run(1, seconds / 2)
    torch.set_num_threads(cores)
    what = 'unmodified reference operators (oracle/_ref), torch CPU fp32 autograd' if kind == 'reference' \
        else 'oracle port of the reference operators (oracle/ops.py), torch CPU fp32 autograd'
    return {'value': v_all, 'unit': 'Mpixel/s', 'cores': cores, 'kind': kind,
            'sample': '%d steps of %dx3x%dx%d (configured batch %dx3x%dx%d), 6-op chain fwd+L1+bwd; %s' % (
                n_all, B, H, W, wl['B'], wl['H'], wl['W'], what),
            'one_thread': {'value': v_one, 'unit': 'Mpixel/s', 'cores': 1, 'sample': '%d steps of the same sample, torch.set_num_threads(1)' % n_one}}


def fused_scale(B, H, W, dev):
    return torch.full((B,), 1.0 / (B * 3 * H * W), device=dev, dtype=torch.float32)


PLANNER_M = 512           # pairs per GPU: eight lock-step batches of 64, two in flight (planner.beam_search_pipelined)
PLANNER_BATCH = 64


def planner_e2e_run(dev, seed, reps=3):
    """beam_search_batch (= the reference's beam_search on every pair, utils/beam_search.py:196-264) over 512 synthetic pairs
    of 3x128x128 in lock-step batches of 64, two in flight, beam 8, the six global operators, max 6 steps, Nelder-Mead fits
    resident on the device (BASELINE config 3's shape; its 1000 pairs are 16 such batches: --workload c3).  -> (best wall-clock seconds of reps - 1 timed repetitions,
    candidates scored, mean steps of the top sequences); the first repetition warms the kernels up."""
    import t2onet_b200 as T
    from t2onet_b200 import planner
    names = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
    exe = T.Executor(T.default_options()).to(dev)
    img, tgt, _ = make_batch(PLANNER_M, 128, 128, seed, dev)
    chunks = [(img[c:c + PLANNER_BATCH], tgt[c:c + PLANNER_BATCH]) for c in range(0, PLANNER_M, PLANNER_BATCH)]
    best = None
    for rep in range(reps):
        cnt = [0]
        torch.cuda.synchronize()
        t0 = time.time()
        res = sum(planner.beam_search_pipelined(chunks, exe, 8, CHAIN, names, 6, 1e-2, workers=2, counter=cnt, images='top'), [])
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rep > 0 and (best is None or dt < best[0]):
            best = (dt, cnt[0])
    return best[0], best[1], sum(len(r[0][0]) for r in res) / PLANNER_M


GIER_M = 64            # pairs per GPU: four lock-step batches of 16, two in flight
GIER_BATCH = 16


def planner_gier_run(dev, seed, reps=2):
    """BASELINE config 5's planner shape: GIER-shaped inputs (3x256x256 pairs, each with the global mask and two local masks that
    belong to brightness and saturation, preprocess/gen_greedy_seqs_GIER.py:36-62), beam 3 and err 1e-3 as that driver sets them,
    the six global operators, Nelder-Mead; 64 pairs in lock-step batches of 16, two in flight.  -> (seconds, candidates, mean steps)"""
    import t2onet_b200 as T
    from t2onet_b200 import planner
    names = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
    exe = T.Executor(T.default_options()).to(dev)
    img, tgt, _ = make_batch(GIER_M, 256, 256, seed, dev)
    masks, idx = [], []
    for m in range(GIER_M):
        m1 = torch.zeros(1, 1, 256, 256, device=dev); m1[..., 32:128 + 8 * (m % 4), 32:160] = 1.0
        m2 = torch.zeros(1, 1, 256, 256, device=dev); m2[..., 128:, 64 + 8 * (m % 3):] = 1.0
        masks.append([torch.ones(1, 1, 256, 256, device=dev), m1, m2])
        idx.append([-1, 0, 2])
    chunks = [(img[c:c + GIER_BATCH], tgt[c:c + GIER_BATCH], {'masks': masks[c:c + GIER_BATCH], 'mask_op_idx': idx[c:c + GIER_BATCH]})
              for c in range(0, GIER_M, GIER_BATCH)]
    best = None
    for rep in range(reps + 1):
        cnt = [0]
        torch.cuda.synchronize()
        t0 = time.time()
        res = sum(planner.beam_search_pipelined(chunks, exe, 3, CHAIN, names, 6, 1e-3, workers=2, counter=cnt, images='top'), [])
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rep > 0 and (best is None or dt < best[0]):
            best = (dt, cnt[0])
    return best[0], best[1], sum(len(r[0][0]) for r in res) / GIER_M


def planner_gier_line(seconds, candidates, mean_steps, n_gpus):
    return {'workload': '%d GIER-shaped pairs of 3x256x256 per GPU (global mask + 2 local masks), beam 3, ops [0,1,2,3,5,6], max_step 6, '
                        'err 1e-3, Nelder-Mead' % GIER_M, 'n_gpus': n_gpus, 'seconds': seconds, 'pairs_per_s': GIER_M * n_gpus / seconds,
            'candidates': candidates, 'candidates_per_s': candidates / seconds, 'mean_steps': mean_steps,
            'how': 'wall clock around beam_search_pipelined with masks (lock-step batches of 16 pairs, two in flight), best of 2; N > 1: pairs sharded by image, slowest rank\'s time'}


def planner_e2e_line(seconds, candidates, mean_steps, n_gpus):
    pairs = PLANNER_M * n_gpus
    return {'workload': '%d pairs of 3x128x128 per GPU, beam 8, ops [0,1,2,3,5,6], max_step 6, err 1e-2, Nelder-Mead' % PLANNER_M,
            'n_gpus': n_gpus, 'seconds': seconds, 'pairs_per_s': pairs / seconds, 'candidates': candidates,
            'candidates_per_s': candidates / seconds, 'mean_steps': mean_steps,
            'how': 'wall clock around beam_search_pipelined (lock-step batches of 64 pairs, two in flight; host bookkeeping and the copies of the top sequences\' images included), best of 2; N > 1: pairs sharded '
                   'by image, no communication during the search, slowest rank\'s time, candidates summed over ranks'}


def extras(TF, dev, wl):
    """Secondary measurements printed with the headline line (rank 0, N=1 only)."""
    ex = {}
    peak, _ = measured_peak_hbm()

    def bench(fn, iters, sync=torch.cuda.synchronize):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync(); s.record()
        for _ in range(iters):
            fn()
        e.record(); sync()
        return s.elapsed_time(e) / iters

    # ---- the Actor's call site end to end (BASELINE config 2 as the reference runs it, models/actor.py:156-170): 5 decoding
    # steps, every row its own operator, parameters from the operators' FC heads (512 -> 512 -> n), L1 to a target, backward
    # down to the FC weights -- (a) the reference's divide_op_group loop over Executor.execute, (b) Executor.execute_rows
    try:
        import t2onet_b200 as T
        torch.manual_seed(10)
        exe = T.Executor(T.default_options()).to(dev)
        B, H, W, STEPS = 64, 128, 128, 5
        gen = torch.Generator().manual_seed(10 + 7000)
        img = torch.rand(B, 3, H, W, generator=gen).to(dev)
        tgt = torch.rand(B, 3, H, W, generator=gen).to(dev)
        feats = [torch.randn(B, 512, generator=gen).to(dev) for _ in range(STEPS)]
        ops_cpu = [torch.tensor([CHAIN[int(v)] for v in torch.randint(0, len(CHAIN), (B,), generator=gen)]) for _ in range(STEPS)]
        ops_dev = [o.to(dev) for o in ops_cpu]

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

        def episode(rows, batched=False):
            exe.zero_grad(set_to_none=True)
            x = img
            for k in range(STEPS):
                x = exe.execute_rows(x, ops_dev[k], None, feats[k], batched_heads=batched)[0] if rows else grouped_step(x, ops_cpu[k], feats[k])
            loss = (x - tgt).abs().mean()
            loss.backward()
            return loss
        t_grp = bench(lambda: episode(False), 10)
        t_row = bench(lambda: episode(True), 10)
        t_bat = bench(lambda: episode(True, True), 10)
        # the same execute_rows episode (device-resident operator ids: no host sync anywhere) captured as ONE CUDA graph
        t_graph, graph_err = None, None
        try:
            # Operator.execute keeps its last parameters (self.param, as the reference does): they hold the eager episodes'
            # autograd graphs -- and their AccumulateGrad nodes, bound to the default stream -- alive; drop them first
            import gc
            for Op in exe.ops:
                Op.param, Op.mask = None, None
            gc.collect()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    episode(True, True)
            torch.cuda.current_stream(dev).wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            exe.zero_grad(set_to_none=True)
            with torch.cuda.graph(gr, capture_error_mode='thread_local'):
                loss_g = episode(True, True)
            t_graph = bench(lambda: gr.replay(), 10)
            ref_loss = episode(True, True).item()
            gr.replay()
            torch.cuda.synchronize()
            assert abs(loss_g.item() - ref_loss) <= 1e-5, (loss_g.item(), ref_loss)
        except Exception as exc:
            graph_err = repr(exc)
        px = B * H * W
        ex['actor_call_site'] = {'workload': '%dx3x%dx%d, %d decoding steps, a random global operator per row and step, FC heads included, '
                                             'L1 + backward to the FC weights' % (B, H, W, STEPS),
                                 'grouped_loop_ms': t_grp, 'execute_rows_ms': t_row, 'speedup': t_grp / t_row,
                                 'execute_rows_batched_heads_ms': t_bat,
                                 'execute_rows_batched_heads_cuda_graph_ms': t_graph, 'cuda_graph_error': graph_err,
                                 'execute_rows_Mpixel_steps_per_s': px * STEPS / t_row / 1e3,
                                 'how': 'eager PyTorch autograd around the kernels, CUDA events, 10 repetitions'}
    except Exception as exc:
        ex['actor_call_site'] = {'error': repr(exc)}
    # ---- C4: the roofline configuration (inputs >> L2)
    try:
        B, H, W = 16, 2048, 3072
        img, tgt, params = make_batch(B, H, W, 10 + 4000, dev)
        px = B * H * W
        t_fused = bench(lambda: TF.chain_forward_backward(img, CHAIN, params, tgt, want_out=True), 5)
        t_fwd = bench(lambda: TF._forward_raw(CHAIN, [0, 1, 2, 3, 27, 35], img, None, 0, torch.cat(params, 1).contiguous(),
                                              36, tgt, True, True, 8), 5)
        t_point = bench(lambda: TF.chain_forward_backward(img, CHAIN[:5], params[:5], tgt, want_out=True), 5)
        ex['c4'] = {
            'workload': WORKLOADS['c4']['desc'],
            'fused_fwd_bwd': {'ms': t_fused, 'Mpixel_per_s': px / t_fused / 1e3, 'GBps_at_36B_px': 36 * px / t_fused / 1e6,
                              'frac_of_measured_peak': 36 * px / t_fused / 1e6 / peak},
            'forward_l1_only': {'ms': t_fwd, 'Mpixel_per_s': px / t_fwd / 1e3, 'GBps_at_36B_px': 36 * px / t_fwd / 1e6,
                                'frac_of_measured_peak': 36 * px / t_fwd / 1e6 / peak},
            'fused_fwd_bwd_pointwise5': {'ms': t_point, 'Mpixel_per_s': px / t_point / 1e3,
                                         'GBps_at_36B_px': 36 * px / t_point / 1e6},
        }
        del img, tgt
        torch.cuda.empty_cache()
    except Exception as exc:   # report, never hide
        ex['c4'] = {'error': repr(exc)}
    # ---- single operators at the C4 shape: Executor.execute forward (24 B/px) and its autograd backward (36 B/px:
    # img + grad_out read, grad_img written) -- the launches the Actor actually issues, one operator per decoding step
    try:
        B, H, W = 16, 2048, 3072
        px = B * H * W
        gen = torch.Generator().manual_seed(10 + 5000)
        dgen = torch.Generator(device=dev).manual_seed(10 + 5000)
        img = torch.rand(B, 3, H, W, generator=dgen, device=dev)
        gout = torch.randn(B, 3, H, W, generator=dgen, device=dev)
        ops_tab = {}
        names = {0: 'brightness', 1: 'contrast', 2: 'saturation', 3: 'color', 5: 'tone', 6: 'sharpness', 8: 'exposure', 9: 'whitebalance',
                 10: 'black&white', 11: 'blur', 12: 'hue'}
        prm = dict(zip(CHAIN, make_params(B, gen, dev)))
        prm[8] = torch.rand(B, 1, generator=gen).to(dev) * 2 - 1
        prm[9] = 0.4 + 1.4 * torch.rand(B, 3, generator=gen).to(dev)
        prm[10] = torch.rand(B, 1, generator=gen).to(dev)
        prm[11] = torch.rand(B, 1, generator=gen).to(dev)
        prm[12] = torch.rand(B, 1, generator=gen).to(dev) * 6.28
        for op, name in names.items():
            p = prm[op].contiguous()
            n = p.shape[1]
            t_f = bench(lambda: TF._forward_raw([op], [0], img, None, 0, p, n, None, True, False, 8), 5)
            t_b = bench(lambda: TF._backward_raw([op], [0], img, None, 0, p, n, gout, None, None, True, False, False, 8), 5)
            ops_tab[name] = {'fwd_ms': t_f, 'fwd_GBps_at_24B_px': 24 * px / t_f / 1e6, 'fwd_frac_of_measured_peak': 24 * px / t_f / 1e6 / peak,
                             'bwd_ms': t_b, 'bwd_GBps_at_36B_px': 36 * px / t_b / 1e6, 'bwd_frac_of_measured_peak': 36 * px / t_b / 1e6 / peak}
        ex['c4_single_ops'] = ops_tab
        # evaluation metrics at the same shape: SSIM (utils/ssim) and L1, 24 B/px each (two images read, one scalar out)
        from t2onet_b200 import metrics as MT
        t_s = bench(lambda: MT.ssim_sum(img, gout), 5)
        t_l = bench(lambda: TF.l1_sum(img, gout), 5)
        ex['c4_metrics'] = {'ssim_ms': t_s, 'ssim_GBps_at_24B_px': 24 * px / t_s / 1e6, 'ssim_frac_of_measured_peak': 24 * px / t_s / 1e6 / peak,
                            'l1_ms': t_l, 'l1_GBps_at_24B_px': 24 * px / t_l / 1e6, 'l1_frac_of_measured_peak': 24 * px / t_l / 1e6 / peak}
        del img, gout
        torch.cuda.empty_cache()
    except Exception as exc:
        ex['c4_single_ops'] = {'error': repr(exc)}
    # ---- per-row operator steps (Actor call sites models/actor.py:156-170): every row its own operator
    try:
        import random
        rnd = random.Random(10 + 6000)
        for tag, (B, H, W) in (('c2', (64, 128, 128)), ('b64_512', (64, 512, 512))):
            px = B * H * W
            gen = torch.Generator().manual_seed(10 + 6000)
            img = torch.rand(B, 3, H, W, generator=gen).to(dev)
            tgt = torch.rand(B, 3, H, W, generator=gen).to(dev)
            gout = torch.randn(B, 3, H, W, generator=gen).to(dev)
            plist = dict(zip(CHAIN, make_params(B, gen, 'cpu')))
            res, scale = {}, fused_scale(B, H, W, dev)
            for K in (1, 5):
                rows = [rnd.sample(CHAIN, K) for _ in range(B)]
                params = torch.zeros(B, K * 24)
                for b in range(B):
                    for k, op in enumerate(rows[b]):
                        v = plist[op][b]
                        params[b, k * 24:k * 24 + v.numel()] = v
                params = params.to(dev)
                ops_dev, ops_host = TF._prep_row_ops(rows, B, torch.device(dev))
                t_f = bench(lambda: TF._rows_forward_raw(ops_dev, ops_host, img, None, 0, params, None, True, False, 8), 10)
                t_b = bench(lambda: TF._rows_backward_raw(ops_dev, ops_host, img, None, 0, params, gout, None, None, True, False, False, 8), 10)
                t_s = bench(lambda: TF._rows_backward_raw(ops_dev, ops_host, img, None, 0, params, None, tgt, scale, False, True, True, 8), 10)
                res['K%d' % K] = {'forward_ms': t_f, 'forward_GBps_at_24B_px': 24 * px / t_f / 1e6,
                                  'backward_ms': t_b, 'backward_GBps_at_36B_px': 36 * px / t_b / 1e6,
                                  'fused_step_ms': t_s, 'fused_step_Mpixel_per_s': px / t_s / 1e3}
            ex['per_row_ops_' + tag] = {'workload': '%dx3x%dx%d, a random operator (K=1) / a random 5-operator order (K=5) per row' % (B, H, W), **res}
    except Exception as exc:
        ex['per_row_ops'] = {'error': repr(exc)}
    # ---- planner candidate scoring: 64 images x 8 states x 168 candidates per state (SURVEY.md section 8d, C3 sweep)
    try:
        S, H, W = 512, 128, 128
        gen = torch.Generator().manual_seed(10 + 3000)
        states = torch.rand(S, 3, H, W, generator=gen).to(dev)
        targets = torch.rand(64, 3, H, W, generator=gen).to(dev)
        ops, prm = [], []
        for op, cnt in ((0, 10), (1, 10), (2, 10), (6, 10), (5, 64), (3, 64)):
            n = {3: 24, 5: 8}.get(op, 1)
            for _ in range(cnt):
                ops.append(op)
                row = torch.zeros(24)
                row[:n] = (0.5 + torch.rand(n, generator=gen)) if n > 1 else torch.rand(1, generator=gen) * 0.5
                prm.append(row)
        per = len(ops)
        cand_state = [s for s in range(S) for _ in range(per)]
        cand_op = ops * S
        cand_param = torch.stack(prm).repeat(S, 1)
        st_t = [s // 8 for s in range(S)]
        cb = TF.CandidateBatch(S, cand_state, cand_op, cand_param, dev, st_t)   # staged once; time only the launches
        t_sc = bench(lambda: TF.score_prepared(states, targets, cb), 5)
        C = S * per
        ex['planner_scoring'] = {'workload': '512 states (64 images x beam 8) of 3x128x128, %d candidates per state' % per,
                                 'candidates': C, 'ms': t_sc, 'candidates_per_s': C / t_sc * 1e3,
                                 'Gpixel_candidates_per_s': C * H * W / t_sc / 1e6}
    except Exception as exc:
        ex['planner_scoring'] = {'error': repr(exc)}
    # ---- the planner end to end through its public API (see planner_e2e_run)
    try:
        ex['planner_e2e'] = planner_e2e_line(*planner_e2e_run(dev, 3010), n_gpus=1)
    except Exception as exc:
        ex['planner_e2e'] = {'error': repr(exc)}
    try:
        ex['planner_gier'] = planner_gier_line(*planner_gier_run(dev, 5010), n_gpus=1)
    except Exception as exc:
        ex['planner_gier'] = {'error': repr(exc)}
    return ex


def run_planner_c3(args):
    """BASELINE config 3 in full: greedy/beam operation planning (the loop of preprocess/gen_greedy_seqs_FiveK.py) over 1000
    synthetic 3x128x128 image pairs, beam width 8, the six global operators, Nelder-Mead -- image-sharded over the ranks,
    64 pairs in lock-step per beam_search_batch call.  `python bench.py --workload c3 [--gpus N]`; one JSON line."""
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    import t2onet_b200 as T
    from t2onet_b200 import planner
    names = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
    exe = T.Executor(T.default_options()).to(dev)
    NPAIRS, BATCH, WORKERS = 1000, int(os.environ.get('T2O_PLANNER_BATCH', 64)), int(os.environ.get('T2O_PLANNER_WORKERS', 2))
    mine = list(range(rank, NPAIRS, world))
    # warm-up: two full batches through the pipelined driver (kernels, the allocator's pools of both worker streams)
    planner.beam_search_pipelined([make_batch(BATCH, 128, 128, 2000 + k, dev)[:2] for k in range(2)], exe, 8, CHAIN, names, 6, 1e-2,
                                  workers=WORKERS, images='top')
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    # the synthetic pairs are generated (and resident in HBM) before the clock starts, like the inputs of every other `value`
    data = []
    for c0 in range(0, len(mine), BATCH):
        idx = mine[c0:c0 + BATCH]
        data.append(make_batch(len(idx), 128, 128, 3010 + 7 * idx[0], dev)[:2])
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.time()
    cnt, steps, per_batch = [0], 0, []
    # the dataset loop, two batches in flight (planner.beam_search_pipelined: one batch's host bookkeeping overlaps the other's
    # device fits; every result equals the sequential call's)
    marks = [time.time()]

    def batches():
        for img, tgt in data:
            marks.append(time.time())
            yield img, tgt
    for res in planner.beam_search_pipelined(batches(), exe, 8, CHAIN, names, 6, 1e-2, workers=WORKERS, counter=cnt, images='top'):
        steps += sum(len(r[0][0]) for r in res)
    per_batch = [round(b - a, 2) for a, b in zip(marks[:-1], marks[1:])]      # (hand-over times of the batches)
    torch.cuda.synchronize()
    t = torch.tensor([time.time() - t0, float(cnt[0]), float(steps)], device=dev, dtype=torch.float64)
    if dist is not None:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t = torch.stack([tmax[0], tsum[1], tsum[2]])
    if rank == 0:
        sec, cand = t[0].item(), int(t[1].item())
        print(json.dumps({'metric': 'planner candidates/s', 'value': cand / sec, 'unit': 'candidates/s', 'n_gpus': world,
                          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'seconds': sec, 'pairs_per_s': NPAIRS / sec, 'rank0_seconds_per_batch': per_batch, 'candidates': cand, 'mean_steps': t[2].item() / NPAIRS,
                          'config': {'workload': 'C3: 1000 pairs of 3x128x128, beam 8, ops [0,1,2,3,5,6], max_step 6, err 1e-2, '
                                                 'Nelder-Mead; image-sharded, %d pairs in lock-step per call, %d calls in flight; the top sequence\'s images copied to the host '
                                                 '(what the dataset driver writes, gen_greedy_seqs_FiveK.py:86-88)' % (BATCH, WORKERS)}}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS) + ['c3'])
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    args = ap.parse_args()
    if args.workload == 'c3':
        return run_planner_c3(args)
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
