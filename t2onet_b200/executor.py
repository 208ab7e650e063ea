"""Drop-in `Executor` (executors/executor.py:14-63 of the reference): same operator registry, index
order, attribute names (checkpoint keys `executor.<name>_op.fc1/fc2.*`) and `execute` signature."""
import torch
import torch.nn as nn

from . import functional as TF
from .operators import (BrightnessOperator, SharpnessOperator, ContrastOperator, InpaintOperator, WhiteOperator,
                        SaturationOperator, ToneOperator, ColorOperator)


class Executor(nn.Module):
    def __init__(self, opt):
        super(Executor, self).__init__()
        self.opt = opt
        self._register_operators(opt)
        self.name_list = [op.short_name for op in self.ops]

    def _register_operators(self, opt):
        # construction order = the reference's (executors/executor.py:21-29): it fixes the RNG stream of
        # the fc initialisation and the state_dict key order
        self.brightness_op = BrightnessOperator(opt)
        self.sharpness_op = SharpnessOperator(opt)
        self.color_op = ColorOperator(opt)
        self.contrast_op = ContrastOperator(opt)
        self.inpaint_op = InpaintOperator(opt)
        self.white_op = WhiteOperator(opt)
        self.saturation_op = SaturationOperator(opt)
        self.tone_op = ToneOperator(opt)
        self.ops = [self.brightness_op, self.contrast_op, self.saturation_op, self.color_op, self.inpaint_op,
                    self.tone_op, self.sharpness_op, self.white_op]

    def execute(self, img, op_ind, mask, features=None, specified_param=None, has_noise=False):
        """execute ONE operator over the batch (executors/executor.py:33-55)
        :param img: (bs, 3, h, w)   :param op_ind: int   :param mask: (bs, 1|3, h, w) or None
        :return out (bs, 3, h, w), param (bs, param_len)"""
        if op_ind < 0:
            bs = img.shape[0]
            return img, torch.zeros(bs, 24, dtype=torch.float).to(img.device)
        Op = self.ops[op_ind]
        if specified_param is not None:
            out = Op.execute(img, mask=mask, features=None, specified_param=specified_param, has_noise=has_noise)
        else:
            out = Op.execute(img, mask=mask, features=features, has_noise=has_noise)
        return out, Op.param

    def execute_chain(self, img, op_inds, params, mask=None):
        """Extension: a whole operator sequence with known parameters in ONE pass over HBM
        (what the planner replays and what BASELINE configs 1/4 measure)."""
        return TF.chain(img, list(op_inds), list(params), mask, getattr(self.opt, 'curve_steps', 8))

    def _row_param(self, Op, features, specified_param, has_noise, n_rows, device):
        """What Operator.execute does to obtain its parameters (models/operators.py:114-125), zero-padded to 24 columns
        like the Actor pads them (models/actor.py:166)."""
        param = Op.extract_parameters(features) if features is not None else specified_param[:, :Op.num_op_param]
        if has_noise:
            param = torch.clamp(param + Op.get_param_noise(n_rows).to(device), Op.lb, Op.ub)
        param = param.float().to(device)
        return torch.nn.functional.pad(param, (0, TF.PARAM_SLOT - param.shape[1]))

    def _batched_head_params(self, features, ops_t):
        """All operators' FC heads (fc1 -> LeakyReLU -> fc2 -> regressor, models/operators.py:73-88) as TWO batched GEMMs
        over the stacked weights instead of two small GEMMs + an activation per operator; every row then keeps the
        parameters of its own operator.  Same arithmetic per head up to the GEMM's summation order (~1e-6)."""
        heads = [(ind, Op) for ind, Op in enumerate(self.ops) if not isinstance(Op, InpaintOperator)]
        G, bs = len(heads), features.shape[0]
        slot = TF.PARAM_SLOT
        W1 = torch.stack([Op.fc1.weight for _, Op in heads])                                   # (G, fc, 2*hidden)
        b1 = torch.stack([Op.fc1.bias for _, Op in heads])
        h = torch.baddbmm(b1.unsqueeze(1), features.unsqueeze(0).expand(G, bs, features.shape[1]), W1.transpose(1, 2))
        h = torch.nn.functional.leaky_relu(h, heads[0][1].lrelu.negative_slope)
        W2 = torch.stack([torch.nn.functional.pad(Op.fc2.weight, (0, 0, 0, slot - Op.num_op_param)) for _, Op in heads])
        b2 = torch.stack([torch.nn.functional.pad(Op.fc2.bias, (0, slot - Op.num_op_param)) for _, Op in heads])
        y = torch.baddbmm(b2.unsqueeze(1), h, W2.transpose(1, 2))                              # (G, bs, 24)
        params = torch.zeros(bs, slot, device=features.device)
        for g, (ind, Op) in enumerate(heads):
            p = torch.nn.functional.pad(Op.op_param_regressor(y[g, :, :Op.num_op_param]), (0, slot - Op.num_op_param))
            params = torch.where((ops_t == ind).view(bs, 1), p, params)
        return params

    def execute_rows(self, img, op_inds, mask, features=None, specified_param=None, has_noise=False, batched_heads=False):
        """Extension: ONE operator step for a batch whose rows use DIFFERENT operators
        -- the Actor's divide_op_group loop (models/actor.py:100-114, 156-170, 245-259) as a single call:
            out, param = executor.execute_rows(img_x, pred_op.view(-1) - 3, mask, context)
        :param op_inds: (bs,) Executor index per row (-1 = <END>: the row passes through, its param row is zeros,
                        executors/executor.py:44-46).  A list / CPU tensor groups the rows per operator for the FC
                        heads exactly like the reference; a CUDA tensor keeps the whole step free of host syncs
                        (every operator head runs on the full batch and the rows select theirs).
        :param features: (bs, 2*hidden) or None   :param specified_param: (bs, >= n) zero-padded rows or None
        :param batched_heads: with device-resident op_inds and features: run all FC heads as two batched GEMMs
                        (_batched_head_params) instead of one pair of small GEMMs per operator
        :return out (bs, 3, h, w), param (bs, 24) zero-padded -- both differentiable."""
        assert (features is None) ^ (specified_param is None)
        bs, dev = img.shape[0], img.device
        on_device = isinstance(op_inds, torch.Tensor) and op_inds.is_cuda
        ops_t = op_inds.view(-1) if isinstance(op_inds, torch.Tensor) else torch.as_tensor(op_inds).view(-1)
        assert ops_t.numel() == bs
        if on_device and batched_heads and features is not None and not has_noise:
            params = self._batched_head_params(features, ops_t)
        elif on_device:
            params = torch.zeros(bs, TF.PARAM_SLOT, device=dev)
            for ind, Op in enumerate(self.ops):
                if isinstance(Op, InpaintOperator):
                    continue
                p = self._row_param(Op, features, specified_param, has_noise, bs, dev)
                params = torch.where((ops_t == ind).view(bs, 1), p, params)
        else:
            ops_l = [max(int(v), -1) for v in ops_t.tolist()]      # op_ind < 0: identity (executors/executor.py:44)
            params = torch.zeros(bs, TF.PARAM_SLOT, device=dev)
            for ind in sorted(set(ops_l)):
                if ind < 0:
                    continue
                Op = self.ops[ind]
                if isinstance(Op, InpaintOperator):
                    raise NotImplementedError('InpaintOperator (EdgeConnect) is outside the B200 hot path')
                rows = torch.tensor([b for b, v in enumerate(ops_l) if v == ind], device=dev)
                f_g = None if features is None else features.index_select(0, rows)
                s_g = None if specified_param is None else specified_param.to(dev).index_select(0, rows)
                params = params.index_copy(0, rows, self._row_param(Op, f_g, s_g, has_noise, rows.numel(), dev))
            ops_t = ops_l
        out = TF.execute_rows(img, ops_t, params, mask, getattr(self.opt, 'curve_steps', 8))
        return out, params

    def get_param_bnd(self, op_ind):
        return self.ops[op_ind].get_param_range()

    def get_param_num(self, op_ind):
        return self.ops[op_ind].num_op_param
