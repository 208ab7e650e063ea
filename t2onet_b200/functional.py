"""Tensor-level API over the C-ABI: fused operator chains with autograd, L1, candidate scoring.

Mirrors the call pattern of the reference (file:line in /root/reference):
  chain(...)           K x Executor.execute(img, op, mask, specified_param=p)   executors/executor.py:33-55
  chain_l1(...)        ... followed by the L1 to a target                        utils/beam_search.py:170-173,
                                                                                 experiments/t2onet/train_seq2seqL1.py:85
  l1_sum / get_dist    (x1 - x2).norm(1) / numel                                 utils/beam_search.py:170-173
  score_candidates     the evaluations inside get_param_naive / beam_search      utils/beam_search.py:77-87,229-237
All of them run the hand-written sm_100a kernels; there is no eager fallback.
"""
import ctypes
import os

import torch

from . import _lib

OP_IDENTITY, OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_COLOR, OP_INPAINT = -1, 0, 1, 2, 3, 4
OP_TONE, OP_SHARPNESS, OP_WHITE, OP_EXPOSURE, OP_WHITEBALANCE = 5, 6, 7, 8, 9
OP_BNW, OP_BLUR, OP_HUE = 10, 11, 12          # classes without an Executor slot (models/operators.py:298, 373, 414)
STENCIL_OPS = (OP_SHARPNESS, OP_BLUR)
CURVE_STEPS = 8


def num_params(op_id, curve_steps=CURVE_STEPS):
    if op_id == OP_COLOR:
        return 3 * curve_steps
    if op_id == OP_TONE:
        return curve_steps
    if op_id == OP_WHITEBALANCE:
        return 3
    if op_id == OP_IDENTITY:
        return 0
    return 1


def _prep_img(t, name):
    if t is None:
        return None
    _lib.require_cuda(t)
    if t.dim() != 4 or t.shape[1] != 3:
        raise _lib.T2OError('%s must be (B, 3, H, W), got %s' % (name, tuple(t.shape)))
    return t.contiguous()


def _prep_mask(mask, img):
    if mask is None:
        return None, 0
    _lib.require_cuda(mask)
    if mask.dim() != 4 or mask.shape[1] not in (1, 3) or mask.shape[0] != img.shape[0] or mask.shape[2:] != img.shape[2:]:
        raise _lib.T2OError('mask must be (B, 1|3, H, W) matching the image, got %s' % (tuple(mask.shape),))
    return mask.contiguous(), mask.shape[1]


def pack_params(op_ids, params, B, device, curve_steps=CURVE_STEPS):
    """List of per-op (B, n_k) tensors -> one (B, sum n_k) row-major tensor + column offsets."""
    offs, cols, off = [], [], 0
    for op, p in zip(op_ids, params):
        n = num_params(op, curve_steps)
        offs.append(off)
        if n == 0:
            continue
        if p.dim() != 2 or p.shape[0] != B or p.shape[1] < n:
            raise _lib.T2OError('operator %d needs a (B=%d, >=%d) parameter tensor, got %s' % (op, B, n, tuple(p.shape)))
        cols.append(p[:, :n])
        off += n
    if off == 0:
        return torch.zeros(B, 1, device=device), offs, 1
    packed = torch.cat(cols, dim=1).contiguous().float()
    return packed, offs, off


def split_segments(op_ids):
    """Launch segments: at most MAX_CHAIN operators and each operator type at most once (the backward kernel
    keeps one register accumulator slot per operator type; the stencil of a second sharpness needs a new pass)."""
    segs, cur, seen = [], [], set()
    for i, op in enumerate(op_ids):
        key = 'stencil' if op in STENCIL_OPS else op          # sharpness and blur share the launch's one stencil
        if len(cur) == _lib.MAX_CHAIN or key in seen:
            segs.append(cur)
            cur, seen = [], set()
        cur.append(i)
        if op >= 0:
            seen.add(key)
    if cur:
        segs.append(cur)
    return segs


def _forward_raw(op_ids, offs, img, mask, mask_ch, packed, pstride, target, want_out, want_l1, curve_steps, flags=0):
    B, _, H, W = img.shape
    lib = _lib.lib()
    out = torch.empty_like(img) if want_out else None
    l1 = torch.empty(B, device=img.device, dtype=torch.float32) if want_l1 else None
    nbytes = lib.t2o_workspace_bytes(B, H, W, pstride)
    ws = _lib.workspace(img.device, nbytes)
    st = lib.t2o_chain_forward(len(op_ids), _lib.int_array(op_ids), _lib.int_array(offs), _lib.ptr(img), _lib.ptr(mask),
                               mask_ch, _lib.ptr(packed), pstride, _lib.ptr(target), _lib.ptr(out), _lib.ptr(l1),
                               B, H, W, curve_steps, flags, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(img.device))
    _lib.check(st)
    return out, l1


def _backward_raw(op_ids, offs, img, mask, mask_ch, packed, pstride, grad_out, target, grad_l1, want_gimg,
                  want_out, want_l1, curve_steps):
    B, _, H, W = img.shape
    lib = _lib.lib()
    gp = torch.empty(B, pstride, device=img.device, dtype=torch.float32)
    gi = torch.empty_like(img) if want_gimg else None
    out = torch.empty_like(img) if want_out else None
    l1 = torch.empty(B, device=img.device, dtype=torch.float32) if want_l1 else None
    nbytes = lib.t2o_workspace_bytes(B, H, W, pstride)
    ws = _lib.workspace(img.device, nbytes)
    st = lib.t2o_chain_backward(len(op_ids), _lib.int_array(op_ids), _lib.int_array(offs), _lib.ptr(img), _lib.ptr(mask),
                                mask_ch, _lib.ptr(packed), pstride, _lib.ptr(grad_out), _lib.ptr(target),
                                _lib.ptr(grad_l1), _lib.ptr(gp), _lib.ptr(gi), _lib.ptr(out), _lib.ptr(l1),
                                B, H, W, curve_steps, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(img.device))
    _lib.check(st)
    return gp, gi, out, l1


class _ChainFn(torch.autograd.Function):
    """One launch segment: out = chain(img; params).  Backward recomputes the chain in-kernel."""

    @staticmethod
    def forward(ctx, img, packed, mask, op_ids, offs, curve_steps):
        mask_c, mask_ch = _prep_mask(mask, img)
        out, _ = _forward_raw(op_ids, offs, img, mask_c, mask_ch, packed, packed.shape[1], None, True, False, curve_steps)
        ctx.save_for_backward(img, packed, mask_c if mask_c is not None else torch.empty(0, device=img.device))
        ctx.meta = (tuple(op_ids), tuple(offs), mask_ch, curve_steps)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        img, packed, mask_c = ctx.saved_tensors
        op_ids, offs, mask_ch, curve_steps = ctx.meta
        mask_c = mask_c if mask_ch else None
        need_img = ctx.needs_input_grad[0]
        gp, gi, _, _ = _backward_raw(op_ids, offs, img, mask_c, mask_ch, packed, packed.shape[1], grad_out.contiguous(),
                                     None, None, need_img, False, False, curve_steps)
        return gi, (gp if ctx.needs_input_grad[1] else None), None, None, None, None


class _ChainL1Fn(torch.autograd.Function):
    """One launch segment ending in the fused L1: l1_sum[b] = sum |chain(img)[b] - target[b]|."""

    @staticmethod
    def forward(ctx, img, packed, target, mask, op_ids, offs, curve_steps):
        mask_c, mask_ch = _prep_mask(mask, img)
        _, l1 = _forward_raw(op_ids, offs, img, mask_c, mask_ch, packed, packed.shape[1], target, False, True, curve_steps)
        ctx.save_for_backward(img, packed, target, mask_c if mask_c is not None else torch.empty(0, device=img.device))
        ctx.meta = (tuple(op_ids), tuple(offs), mask_ch, curve_steps)
        return l1

    @staticmethod
    def backward(ctx, grad_l1):
        img, packed, target, mask_c = ctx.saved_tensors
        op_ids, offs, mask_ch, curve_steps = ctx.meta
        mask_c = mask_c if mask_ch else None
        need_img = ctx.needs_input_grad[0]
        gp, gi, _, _ = _backward_raw(op_ids, offs, img, mask_c, mask_ch, packed, packed.shape[1], None, target,
                                     grad_l1.contiguous().float(), need_img, False, False, curve_steps)
        return gi, (gp if ctx.needs_input_grad[1] else None), None, None, None, None, None


def _seg_inputs(op_ids, params, seg, B, device, curve_steps):
    ops = [op_ids[i] for i in seg]
    packed, offs, _ = pack_params(ops, [params[i] for i in seg], B, device, curve_steps)
    return ops, packed, offs


def chain(img, op_ids, params, mask=None, curve_steps=CURVE_STEPS):
    """Apply operators op_ids[k] with parameters params[k] ((B, n_k) tensors) in sequence.
    Differentiable w.r.t. img and params.  Identity steps (op id < 0) pass the image through."""
    img = _prep_img(img, 'img')
    op_ids = [int(o) for o in op_ids]
    if all(o < 0 for o in op_ids):
        return img
    keep = [i for i, o in enumerate(op_ids) if o >= 0]
    op_ids, params = [op_ids[i] for i in keep], [params[i] for i in keep]
    B = img.shape[0]
    for seg in split_segments(op_ids):
        ops, packed, offs = _seg_inputs(op_ids, params, seg, B, img.device, curve_steps)
        img = _ChainFn.apply(img, packed, mask, ops, offs, curve_steps)
    return img


def chain_l1(img, op_ids, params, target, mask=None, curve_steps=CURVE_STEPS):
    """Per-image sum |chain(img) - target| (B,), without materialising the edited image.
    Differentiable w.r.t. img and params."""
    img, target = _prep_img(img, 'img'), _prep_img(target, 'target')
    op_ids = [int(o) for o in op_ids]
    keep = [i for i, o in enumerate(op_ids) if o >= 0]
    op_ids, params = [op_ids[i] for i in keep], [params[i] for i in keep]
    if not op_ids:
        return l1_sum(img, target)
    B = img.shape[0]
    segs = split_segments(op_ids)
    for seg in segs[:-1]:
        ops, packed, offs = _seg_inputs(op_ids, params, seg, B, img.device, curve_steps)
        img = _ChainFn.apply(img, packed, mask, ops, offs, curve_steps)
    ops, packed, offs = _seg_inputs(op_ids, params, segs[-1], B, img.device, curve_steps)
    return _ChainL1Fn.apply(img, packed, target, mask, ops, offs, curve_steps)


def chain_forward_backward(img, op_ids, params, target, mask=None, want_out=True, want_grad_img=False,
                           loss_scale=None, curve_steps=CURVE_STEPS):
    """ONE kernel launch for a whole training-style step on a single segment: edited image, per-image L1
    sums to `target`, and the gradients of  loss = sum_b loss_scale[b] * l1_sum[b]  w.r.t. every
    operator parameter (and optionally the input image).  loss_scale defaults to 1/numel (the mean L1 of
    experiments/t2onet/train_seq2seqL1.py:85).  Returns (out, l1_sum, [grad_param_k], grad_img)."""
    img, target = _prep_img(img, 'img'), _prep_img(target, 'target')
    op_ids = [int(o) for o in op_ids]
    if len(split_segments(op_ids)) != 1 or any(o < 0 for o in op_ids):
        raise _lib.T2OError('chain_forward_backward takes one launch segment (<= 8 ops, each operator type at most once, no identity)')
    B = img.shape[0]
    packed, offs, pstride = pack_params(op_ids, params, B, img.device, curve_steps)
    mask_c, mask_ch = _prep_mask(mask, img)
    if loss_scale is None:
        loss_scale = torch.full((B,), 1.0 / img.numel(), device=img.device, dtype=torch.float32)
    gp, gi, out, l1 = _backward_raw(op_ids, offs, img, mask_c, mask_ch, packed, pstride, None, target,
                                    loss_scale.contiguous().float(), want_grad_img, want_out, True, curve_steps)
    grads = [gp[:, o:o + num_params(op, curve_steps)] for op, o in zip(op_ids, offs)]
    return out, l1, grads, gi


class FusedStep:
    """A prepared fused training step (one t2o_chain_backward launch) for a fixed chain and batch shape.

    Everything that does not depend on the data -- the host-side operator / offset arrays, the loss scale, the
    workspace and (with reuse_outputs=True) the output buffers -- is set up once, so a call is one ctypes call:
        out, l1_sum, grad_packed, grad_img = step(img, packed_params, target)
    `packed_params` is the (B, sum n_k) row-major parameter table of pack_params(); grad_packed has the same
    layout (step.split(grad_packed) gives the per-operator views).  With reuse_outputs=True the returned tensors
    are overwritten by the next call."""

    def __init__(self, op_ids, B, H, W, device, want_out=True, want_grad_img=False, reuse_outputs=False,
                 curve_steps=CURVE_STEPS):
        self.op_ids = [int(o) for o in op_ids]
        if len(split_segments(self.op_ids)) != 1 or any(o < 0 for o in self.op_ids):
            raise _lib.T2OError('FusedStep takes one launch segment (<= 8 ops, each operator type at most once, no identity)')
        self.B, self.H, self.W, self.device, self.curve_steps = B, H, W, torch.device(device), curve_steps
        self.offs, off = [], 0
        for op in self.op_ids:
            self.offs.append(off)
            off += num_params(op, curve_steps)
        self.pstride = max(off, 1)
        self.want_out, self.want_grad_img, self.reuse = want_out, want_grad_img, reuse_outputs
        self._ops_c, self._offs_c = _lib.int_array(self.op_ids), _lib.int_array(self.offs)
        self._lib = _lib.lib()
        self._ws_bytes = self._lib.t2o_workspace_bytes(B, H, W, self.pstride)
        self.loss_scale = torch.full((B,), 1.0 / (B * 3 * H * W), device=self.device, dtype=torch.float32)
        self._bufs = None

    def _outputs(self):
        if self._bufs is not None:
            return self._bufs
        shape = (self.B, 3, self.H, self.W)
        gp = torch.empty(self.B, self.pstride, device=self.device, dtype=torch.float32)
        gi = torch.empty(shape, device=self.device, dtype=torch.float32) if self.want_grad_img else None
        out = torch.empty(shape, device=self.device, dtype=torch.float32) if self.want_out else None
        l1 = torch.empty(self.B, device=self.device, dtype=torch.float32)
        bufs = (out, l1, gp, gi)
        if self.reuse:
            self._bufs = bufs
        return bufs

    def split(self, packed):
        return [packed[:, o:o + num_params(op, self.curve_steps)] for op, o in zip(self.op_ids, self.offs)]

    def __call__(self, img, packed_params, target, mask=None, loss_scale=None):
        shape = (self.B, 3, self.H, self.W)
        if tuple(img.shape) != shape or tuple(target.shape) != shape or tuple(packed_params.shape) != (self.B, self.pstride):
            raise _lib.T2OError('FusedStep: expected img/target %s and params %s' % (shape, (self.B, self.pstride)))
        _lib.require_cuda(img, target, packed_params)
        if not (img.is_contiguous() and target.is_contiguous() and packed_params.is_contiguous()):
            raise _lib.T2OError('FusedStep: inputs must be contiguous')
        mask_c, mask_ch = _prep_mask(mask, img)
        scale = self.loss_scale if loss_scale is None else loss_scale.contiguous().float()
        out, l1, gp, gi = self._outputs()
        ws = _lib.workspace(self.device, self._ws_bytes)
        st = self._lib.t2o_chain_backward(len(self.op_ids), self._ops_c, self._offs_c, _lib.ptr(img), _lib.ptr(mask_c), mask_ch,
                                          _lib.ptr(packed_params), self.pstride, None, _lib.ptr(target), _lib.ptr(scale),
                                          _lib.ptr(gp), _lib.ptr(gi), _lib.ptr(out), _lib.ptr(l1),
                                          self.B, self.H, self.W, self.curve_steps, _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr(self.device))
        _lib.check(st)
        return out, l1, gp, gi


# ------------------------------------------------------------------------------------------ per-row chains
PARAM_SLOT = _lib.MAX_OP_PARAMS      # the Actor's zero-padded parameter rows (models/actor.py:166)


def _prep_row_ops(row_ops, B, device):
    """-> (device int32 (B, K) tensor, ctypes host copy or None).  A list / CPU tensor is known on the host (validated
    up front, any K); a CUDA tensor stays on the device (K == 1, no host sync)."""
    # every negative id is the identity, as Executor.execute treats op_ind < 0 (executors/executor.py:44): the Actor
    # passes vocabulary id - 3, i.e. -3 / -2 / -1 for <NONE> / <START> / <END> (models/actor.py:146,165)
    if isinstance(row_ops, torch.Tensor) and row_ops.is_cuda:
        t = row_ops.reshape(B, -1).to(torch.int32).clamp_min(OP_IDENTITY).contiguous()
        return t, None
    t = torch.as_tensor(row_ops, dtype=torch.int32).reshape(B, -1).clamp_min(OP_IDENTITY).contiguous()
    host = _lib.int_array(t.flatten().tolist())
    return t.to(device, non_blocking=True), host


def _rows_forward_raw(ops_dev, ops_host, img, mask, mask_ch, params, target, want_out, want_l1, curve_steps):
    B, _, H, W = img.shape
    K, pstride = ops_dev.shape[1], params.shape[1]
    lib = _lib.lib()
    out = torch.empty_like(img) if want_out else None
    l1 = torch.empty(B, device=img.device, dtype=torch.float32) if want_l1 else None
    ws = _lib.workspace(img.device, lib.t2o_workspace_bytes(B, H, W, pstride))
    st = lib.t2o_rows_forward(K, _lib.ptr(ops_dev), ops_host, PARAM_SLOT, _lib.ptr(img), _lib.ptr(mask), mask_ch,
                              _lib.ptr(params), pstride, _lib.ptr(target), _lib.ptr(out), _lib.ptr(l1),
                              _lib.ptr(_lib.status_word(img.device)), B, H, W, curve_steps,
                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(img.device))
    _lib.check(st)
    return out, l1


def _rows_backward_raw(ops_dev, ops_host, img, mask, mask_ch, params, grad_out, target, grad_l1, want_gimg, want_out,
                       want_l1, curve_steps):
    B, _, H, W = img.shape
    K, pstride = ops_dev.shape[1], params.shape[1]
    lib = _lib.lib()
    gp = torch.empty(B, pstride, device=img.device, dtype=torch.float32)
    gi = torch.empty_like(img) if want_gimg else None
    out = torch.empty_like(img) if want_out else None
    l1 = torch.empty(B, device=img.device, dtype=torch.float32) if want_l1 else None
    ws = _lib.workspace(img.device, lib.t2o_workspace_bytes(B, H, W, pstride))
    st = lib.t2o_rows_backward(K, _lib.ptr(ops_dev), ops_host, PARAM_SLOT, _lib.ptr(img), _lib.ptr(mask), mask_ch,
                               _lib.ptr(params), pstride, _lib.ptr(grad_out), _lib.ptr(target), _lib.ptr(grad_l1),
                               _lib.ptr(gp), _lib.ptr(gi), _lib.ptr(out), _lib.ptr(l1),
                               _lib.ptr(_lib.status_word(img.device)), B, H, W, curve_steps,
                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr(img.device))
    _lib.check(st)
    return gp, gi, out, l1


class _RowsFn(torch.autograd.Function):
    """out[b] = chain_b(img[b]; params[b]) with a per-row operator chain.  Backward recomputes in-kernel."""

    @staticmethod
    def forward(ctx, img, params, mask, ops_dev, ops_host, curve_steps):
        mask_c, mask_ch = _prep_mask(mask, img)
        out, _ = _rows_forward_raw(ops_dev, ops_host, img, mask_c, mask_ch, params, None, True, False, curve_steps)
        ctx.save_for_backward(img, params, ops_dev, mask_c if mask_c is not None else torch.empty(0, device=img.device))
        ctx.meta = (ops_host, mask_ch, curve_steps)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        img, params, ops_dev, mask_c = ctx.saved_tensors
        ops_host, mask_ch, curve_steps = ctx.meta
        mask_c = mask_c if mask_ch else None
        gp, gi, _, _ = _rows_backward_raw(ops_dev, ops_host, img, mask_c, mask_ch, params, grad_out.contiguous(), None, None,
                                          ctx.needs_input_grad[0], False, False, curve_steps)
        return gi, (gp if ctx.needs_input_grad[1] else None), None, None, None, None


def _prep_row_params(params, B, K):
    _lib.require_cuda(params)
    if params.dim() != 2 or params.shape[0] != B or params.shape[1] != K * PARAM_SLOT:
        raise _lib.T2OError('per-row parameters must be (B=%d, K*%d=%d), got %s' % (B, PARAM_SLOT, K * PARAM_SLOT, tuple(params.shape)))
    return params.contiguous()


def execute_rows(img, row_ops, params, mask=None, curve_steps=CURVE_STEPS):
    """Every batch row applies its OWN operator chain: row b runs operators row_ops[b, 0..K) (Executor indices,
    -1 = <END> = pass through) with parameters params[b, k*24 : k*24 + n] (zero-padded 24-float slots, the Actor's
    layout).  One launch per tiling over the original tensors -- replaces divide_op_group + index_select + per-group
    Executor.execute + cat + index_select (models/actor.py:100-114, 156-170, 245-259).  Differentiable w.r.t. img
    and params.  row_ops: (B,) / (B, K) list, CPU tensor (validated on the host) or CUDA tensor (K == 1, sync-free)."""
    img = _prep_img(img, 'img')
    B = img.shape[0]
    ops_dev, ops_host = _prep_row_ops(row_ops, B, img.device)
    params = _prep_row_params(params, B, ops_dev.shape[1])
    return _RowsFn.apply(img, params, mask, ops_dev, ops_host, curve_steps)


def rows_forward_backward(img, row_ops, params, target, mask=None, want_out=True, want_grad_img=False, loss_scale=None,
                          curve_steps=CURVE_STEPS):
    """The fused training-style step of chain_forward_backward for per-row chains: edited image, per-image L1 sums and
    the gradients of  sum_b loss_scale[b] * l1_sum[b]  (default: the mean L1) in one launch per tiling.
    Returns (out, l1_sum, grad_params (B, K*24), grad_img)."""
    img, target = _prep_img(img, 'img'), _prep_img(target, 'target')
    B = img.shape[0]
    ops_dev, ops_host = _prep_row_ops(row_ops, B, img.device)
    params = _prep_row_params(params, B, ops_dev.shape[1])
    mask_c, mask_ch = _prep_mask(mask, img)
    if loss_scale is None:
        loss_scale = torch.full((B,), 1.0 / img.numel(), device=img.device, dtype=torch.float32)
    gp, gi, out, l1 = _rows_backward_raw(ops_dev, ops_host, img, mask_c, mask_ch, params, None, target,
                                         loss_scale.contiguous().float(), want_grad_img, want_out, True, curve_steps)
    return out, l1, gp, gi


def rows_status(device, clear=True):
    """Bit 0 set: a per-row kernel met an invalid operator id since the last clear (it treated the row as identity).
    Reading synchronises; only device-resident row_ops (no host copy) can trip it."""
    w = _lib.status_word(torch.device(device))
    v = int(w.item())
    if clear and v:
        w.zero_()
    return v


def process_raw(img, op_id, param, curve_steps=CURVE_STEPS):
    """Operator.process(img, param) itself (no mask blend, no clamp): models/operators.py:128."""
    img = _prep_img(img, 'img')
    packed, offs, pstride = pack_params([op_id], [param], img.shape[0], img.device, curve_steps)
    out, _ = _forward_raw([op_id], offs, img, None, 0, packed, pstride, None, True, False, curve_steps,
                          flags=_lib.FLAG_RAW_PROCESS)
    return out


def l1_sum(a, b):
    """Per-image sum |a - b| -> (B,)."""
    _lib.require_cuda(a, b)
    if a.shape != b.shape:
        raise _lib.T2OError('l1_sum: shape mismatch %s vs %s' % (tuple(a.shape), tuple(b.shape)))
    a, b = a.contiguous(), b.contiguous()
    B = a.shape[0]
    n = a.numel() // B
    lib = _lib.lib()
    out = torch.empty(B, device=a.device, dtype=torch.float32)
    ws = _lib.workspace(a.device, (1 << 18) + B * (n // 4096 + 2) * 4)     # counters + one partial per >= 4096-float tile
    st = lib.t2o_l1_sum(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), B, n, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(a.device))
    _lib.check(st)
    return out


class CandidateBatch:
    """Device-resident candidate list for score_candidates (build once, score many times)."""

    def __init__(self, S, cand_state, cand_op, cand_param, device, state_target=None, masks=None, cand_mask=None):
        cs = torch.as_tensor(cand_state, dtype=torch.int64).cpu()
        self.C = int(cs.numel())
        self.S = S
        if self.C and (bool((cs[1:] < cs[:-1]).any()) or int(cs.min()) < 0 or int(cs.max()) >= S):
            raise _lib.T2OError('cand_state must be ascending and within [0, S)')
        begin = torch.zeros(S + 1, dtype=torch.int64)
        if self.C:
            begin[1:] = torch.cumsum(torch.bincount(cs, minlength=S), 0)
        self.begin_host = begin.to(torch.int32).contiguous()          # (host copy: sizes the resident Nelder-Mead launch)
        self.begin = self.begin_host.to(device)
        self.ops = torch.as_tensor(cand_op, dtype=torch.int32).to(device).contiguous()
        prm = torch.as_tensor(cand_param, dtype=torch.float32)
        if prm.dim() != 2 or prm.shape[0] != self.C or prm.shape[1] > _lib.MAX_OP_PARAMS:
            raise _lib.T2OError('cand_param must be (C, <=24)')
        if prm.shape[1] < _lib.MAX_OP_PARAMS:
            prm = torch.cat([prm, prm.new_zeros(self.C, _lib.MAX_OP_PARAMS - prm.shape[1])], 1)
        self.prm = prm.to(device).contiguous()
        self.state_target = None if state_target is None else \
            torch.as_tensor(state_target, dtype=torch.int32).to(device).contiguous()
        # masks (n_masks, 1|3, H, W) + the mask index of every candidate (-1: edit everywhere), the GIER planner's inputs
        self.masks, self.cand_mask = None, None
        if masks is not None:
            if masks.dim() != 4 or masks.shape[1] not in (1, 3) or not masks.is_cuda:
                raise _lib.T2OError('masks must be a CUDA tensor (n_masks, 1|3, H, W)')
            cm = torch.as_tensor(cand_mask, dtype=torch.int64).cpu()
            if cm.numel() != self.C or (self.C and (int(cm.max()) >= masks.shape[0])):
                raise _lib.T2OError('cand_mask must hold one mask index (< n_masks, or -1) per candidate')
            self.masks = masks.detach().float().contiguous()
            self.cand_mask = cm.to(torch.int32).to(device).contiguous()


def score_prepared(states, targets, cb, curve_steps=CURVE_STEPS):
    """One t2o_score_candidates launch on a prepared CandidateBatch -> (C,) float32 CUDA tensor."""
    dev = states.device
    S, _, H, W = states.shape
    out = torch.empty(cb.C, device=dev, dtype=torch.float32)
    if cb.C == 0:
        return out
    lib = _lib.lib()
    ws = _lib.workspace(dev, lib.t2o_score_workspace_bytes(S, cb.C, H, W))
    if cb.masks is not None:
        if tuple(cb.masks.shape[2:]) != (H, W):
            raise _lib.T2OError('masks must have the states\' height and width')
        st = lib.t2o_score_candidates_masked(_lib.ptr(states), S, _lib.ptr(targets), targets.shape[0], _lib.ptr(cb.state_target),
                                             _lib.ptr(cb.begin), _lib.ptr(cb.ops), _lib.ptr(cb.prm), _lib.ptr(cb.cand_mask),
                                             _lib.ptr(cb.masks), cb.masks.shape[0], cb.masks.shape[1], cb.C, _lib.ptr(out), H, W,
                                             curve_steps, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    else:
        st = lib.t2o_score_candidates(_lib.ptr(states), S, _lib.ptr(targets), targets.shape[0], _lib.ptr(cb.state_target),
                                      _lib.ptr(cb.begin), _lib.ptr(cb.ops), _lib.ptr(cb.prm), cb.C, _lib.ptr(out), H, W,
                                      curve_steps, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(st)
    return out


def score_candidates(states, targets, cand_state, cand_op, cand_param, state_target=None, curve_steps=CURVE_STEPS,
                     masks=None, cand_mask=None):
    """Score C single-operator candidates: l1_sum[c] = sum |clamp(op_c(states[cand_state[c]]; param_c)) - target|.

    states (S,3,H,W), targets (T,3,H,W) on the GPU; cand_state / cand_op: int sequences or tensors (C,),
    cand_state ascending; cand_param (C, <=24) float.  state_target (S,) picks each state's target
    (default s % T).  masks (n_masks, 1|3, H, W) + cand_mask (C,): candidate c edits inside mask cand_mask[c] (-1: everywhere).
    Returns a (C,) float32 CUDA tensor in candidate order."""
    states, targets = _prep_img(states, 'states'), _prep_img(targets, 'targets')
    cb = CandidateBatch(states.shape[0], cand_state, cand_op, cand_param, states.device, state_target, masks, cand_mask)
    return score_prepared(states, targets, cb, curve_steps)


def topk_min(values, seg_begin, k):
    """The k smallest values of every segment in ascending order (ties by the smaller index, NaN last): the argsort + [:beam_size]
    of a planner step (utils/beam_search.py:252-256) for many searches at once, on the device.
    values (C,) float32 CUDA; seg_begin (n_seg + 1,) ints.  -> (idx (n_seg, k) int32 into `values`, -1 padded; val (n_seg, k))."""
    _lib.require_cuda(values)
    values = values.detach().float().contiguous()
    sb = torch.as_tensor(seg_begin, dtype=torch.int32).to(values.device).contiguous()
    n_seg = sb.numel() - 1
    idx = torch.empty(n_seg, k, dtype=torch.int32, device=values.device)
    val = torch.empty(n_seg, k, dtype=torch.float32, device=values.device)
    _lib.check(_lib.lib().t2o_topk_min(_lib.ptr(values), _lib.ptr(sb), n_seg, k, _lib.ptr(idx), _lib.ptr(val), _lib.stream_ptr(values.device)))
    return idx, val


import threading
_CAPTURE_LOCK = threading.Lock()


class DeviceNelderMead:
    """P Nelder-Mead fits that live on the GPU (t2o_nm_start / t2o_nm_advance), each scored as candidate p of
    t2o_score_candidates: argmin_param L1(op(states[prob_state[p]]; param), its target) with scipy's defaults, as
    get_param_naive runs them one by one through scipy (utils/beam_search.py:65-91).

    prob_state must be ascending.  run() enqueues rounds of (score, advance) without host synchronisation -- after a
    warm-up the rounds replay as a CUDA graph -- and looks at the fits' state only every `check_every` rounds.
    Results: x (P, 24) float64 (columns >= n are 0), fun (P,) float64, nit / nfev / status (P,) int32."""

    ROWS = _lib.MAX_OP_PARAMS + 1

    def __init__(self, states, targets, prob_state, prob_op, x0, state_target=None, curve_steps=CURVE_STEPS, numel=None,
                 masks=None, prob_mask=None):
        self.states, self.targets = _prep_img(states, 'states'), _prep_img(targets, 'targets')
        dev = self.states.device
        self.dev, self.L = dev, curve_steps
        S, _, H, W = self.states.shape
        self.S, self.H, self.W = S, H, W
        P = len(prob_op)
        self.P = P
        self.numel = float(numel if numel is not None else 3 * H * W)
        n_dims = [num_params(int(o), curve_steps) for o in prob_op]
        if any(n < 1 or n > _lib.MAX_OP_PARAMS for n in n_dims):
            raise _lib.T2OError('Nelder-Mead fits need operators with 1..24 parameters')
        import numpy as np
        x0h = np.zeros((P, _lib.MAX_OP_PARAMS), dtype=np.float64)
        conv = {}                                                   # (the planner passes one start vector object per operator)
        for i, (v, n) in enumerate(zip(x0, n_dims)):
            a = conv.get(id(v))
            if a is None:
                a = conv[id(v)] = (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)).reshape(-1)
            x0h[i, :n] = a[:n]
        x0m = torch.from_numpy(x0h)
        self.cb = CandidateBatch(S, prob_state, prob_op, torch.zeros(P, _lib.MAX_OP_PARAMS), dev, state_target, masks, prob_mask)
        self.n_dims = torch.tensor(n_dims, dtype=torch.int32, device=dev)
        self.prob_op_host = torch.as_tensor(prob_op, dtype=torch.int32).cpu().contiguous()
        self.prob_op = self.prob_op_host.to(dev)
        self.x0 = x0m.to(dev)
        f64 = dict(dtype=torch.float64, device=dev)
        self.sim = torch.empty(P, self.ROWS, _lib.MAX_OP_PARAMS, **f64)
        self.fsim = torch.empty(P, self.ROWS, **f64)
        self.vec = torch.zeros(P, 3, _lib.MAX_OP_PARAMS, **f64)
        self.fxr = torch.zeros(P, **f64)
        self.xbest = torch.zeros(P, _lib.MAX_OP_PARAMS, **f64)
        self.fbest = torch.zeros(P, **f64)
        self.perm = torch.zeros(P, self.ROWS, dtype=torch.int32, device=dev)
        self.ctl = torch.zeros(P, 8, dtype=torch.int32, device=dev)
        self.l1 = torch.zeros(P, dtype=torch.float32, device=dev)
        self.state = _lib.NMState(*[t.data_ptr() for t in (self.sim, self.fsim, self.vec, self.fxr, self.xbest, self.fbest,
                                                            self.perm, self.ctl)])
        self.rounds = 0
        lib = _lib.lib()
        self.ws = _lib.workspace(dev, lib.t2o_score_workspace_bytes(S, P, H, W))
        _lib.check(lib.t2o_nm_start(ctypes.byref(self.state), P, _lib.ptr(self.n_dims), _lib.ptr(self.prob_op), _lib.ptr(self.x0),
                                    _lib.ptr(self.cb.prm), _lib.ptr(self.cb.ops), _lib.stream_ptr(dev)))

    def _round(self):
        lib, cb = _lib.lib(), self.cb
        sp = _lib.stream_ptr(self.dev)
        if cb.masks is not None:
            _lib.check(lib.t2o_score_candidates_masked(_lib.ptr(self.states), self.S, _lib.ptr(self.targets), self.targets.shape[0],
                                                       _lib.ptr(cb.state_target), _lib.ptr(cb.begin), _lib.ptr(cb.ops), _lib.ptr(cb.prm),
                                                       _lib.ptr(cb.cand_mask), _lib.ptr(cb.masks), cb.masks.shape[0], cb.masks.shape[1],
                                                       self.P, _lib.ptr(self.l1), self.H, self.W, self.L, _lib.ptr(self.ws),
                                                       self.ws.numel(), sp))
        else:
            _lib.check(lib.t2o_score_candidates(_lib.ptr(self.states), self.S, _lib.ptr(self.targets), self.targets.shape[0],
                                                _lib.ptr(cb.state_target), _lib.ptr(cb.begin), _lib.ptr(cb.ops), _lib.ptr(cb.prm),
                                                self.P, _lib.ptr(self.l1), self.H, self.W, self.L, _lib.ptr(self.ws),
                                                self.ws.numel(), sp))
        _lib.check(lib.t2o_nm_advance(ctypes.byref(self.state), self.P, _lib.ptr(self.l1), self.numel, _lib.ptr(cb.prm),
                                      _lib.ptr(cb.ops), sp))

    def active(self):
        """Number of unfinished fits (synchronises)."""
        return int((self.ctl[:, 1] != 6).sum().item())

    def run_resident(self, max_rounds=None):
        """All rounds in one launch (t2o_nm_run_resident): the states and the fits' simplices stay in shared memory for the
        life of the fits.  -> True, or False where the shape is not eligible (the caller runs the rounds)."""
        limit = 200 * _lib.MAX_OP_PARAMS + 8 if max_rounds is None else max_rounds
        lib, cb = _lib.lib(), self.cb
        ws = _lib.workspace(self.dev, max(lib.t2o_score_workspace_bytes(self.S, self.P, self.H, self.W), 1 << 18))
        m = cb.masks
        st = lib.t2o_nm_run_resident(_lib.ptr(self.states), self.S, _lib.ptr(self.targets), self.targets.shape[0], _lib.ptr(cb.state_target),
                                     _lib.ptr(cb.begin), _lib.ptr(cb.cand_mask), _lib.ptr(m), 0 if m is None else m.shape[0],
                                     0 if m is None else m.shape[1], ctypes.byref(self.state), self.P, ctypes.c_float(self.numel),
                                     _lib.ptr(cb.prm), _lib.ptr(cb.ops), cb.begin_host.data_ptr(), self.prob_op_host.data_ptr(),
                                     self.H, self.W, self.L, limit, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self.dev))
        if st == 2:                                                 # T2O_ERR_UNSUPPORTED
            return False
        _lib.check(st)
        self.rounds += limit
        return True

    def run(self, check_every=64, use_graph=True, max_rounds=None):
        if os.environ.get('T2O_NM_RESIDENT', '1') != '0' and self.rounds == 0 and self.run_resident(max_rounds):
            if max_rounds is not None or self.active() == 0:
                return self.result()
        limit = 200 * _lib.MAX_OP_PARAMS + 8 if max_rounds is None else max_rounds
        for _ in range(check_every):                               # eager warm-up (also sets the kernels' attributes)
            self._round()
        self.rounds += check_every
        graph = None
        while self.rounds < limit and self.active() > 0:
            if use_graph and graph is None:
                # the rounds are identical launches over fixed buffers: record them once, replay from now on
                # (capture_begin / capture_end directly: torch.cuda.graph() would also run gc.collect() and empty the
                # allocator cache at every capture -- tens of milliseconds per planner step; nothing is allocated here)
                side = torch.cuda.Stream(self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                graph = torch.cuda.CUDAGraph()
                # (one capture at a time, in thread-local mode: planner.beam_search_pipelined runs batches on several threads, and
                # a global-mode capture forbids the other threads' allocations and copies while it lasts)
                with _CAPTURE_LOCK, torch.cuda.stream(side):       # _round() launches on the capturing stream
                    graph.capture_begin(capture_error_mode='thread_local')
                    try:
                        for _ in range(check_every):
                            self._round()
                    finally:
                        graph.capture_end()
                torch.cuda.current_stream(self.dev).wait_stream(side)
            if graph is not None:
                graph.replay()
            else:
                for _ in range(check_every):
                    self._round()
            self.rounds += check_every
        return self.result()

    def result(self):
        ctl = self.ctl.cpu()
        return {'x': self.xbest.cpu(), 'fun': self.fbest.cpu(), 'nit': ctl[:, 4].clone(), 'nfev': ctl[:, 3].clone(),
                'status': ctl[:, 5].clone(), 'done': (ctl[:, 1] == 6), 'n': ctl[:, 0].clone()}
