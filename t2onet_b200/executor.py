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

    def get_param_bnd(self, op_ind):
        return self.ops[op_ind].get_param_range()

    def get_param_num(self, op_ind):
        return self.ops[op_ind].num_op_param
