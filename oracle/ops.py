"""Operator / executor oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional PyTorch-CPU restatement (fp32, differentiable through autograd) of
the reference's global editing operators.  Every function cites the reference
lines it follows; the torch op sequence is kept the same so that fp32 rounding
and autograd's tie rules (clamp closed interval, binary min/max 0.5/0.5 split,
first-index max/min along channels) agree with the reference.

Operator ids are the reference ``Executor`` indices
(/root/reference/executors/executor.py:30) plus extension ids for the operator
classes that exist in models/operators.py without an Executor slot.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import hsv as _hsv

# executors/executor.py:30  self.ops = [brightness, contrast, saturation, color, inpaint, tone, sharpness, white]
OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_COLOR, OP_INPAINT, OP_TONE, OP_SHARPNESS, OP_WHITE = range(8)
# not registered in the Executor, but named by the north star (models/operators.py:186,527)
OP_EXPOSURE, OP_WHITEBALANCE = 8, 9
# further classes without an Executor slot (models/operators.py:298, 373, 414)
OP_BNW, OP_BLUR, OP_HUE = 10, 11, 12
OP_IDENTITY = -1  # executors/executor.py:44-46

# planner / dataset action names (preprocess/gen_greedy_seqs_FiveK.py:40)
ACTION_NAMES = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']


def default_cfg(**over):
    """Operator options with the reference defaults (options/fiveK_base_options.py:30-54)."""
    cfg = SimpleNamespace(
        hidden_size=256, operator_fc_dim=512, discrete_param=0, discrete_step=10,
        exposure_range=3.5, sharpness_range=1.5, brightness_range=2, curve_steps=8,
        tone_curve_range=(0.5, 2), color_curve_range=(0.90, 1.10), saturation_range=(-0.2, 0.8),
        param_noise_factor=0.0, explore_prob=0.0)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def num_params(op_id, cfg=None):
    """models/operators.py: num_op_param of each class (:190,:228,:263,:336,:458,:498,:532,:563,:599,:629)."""
    L = 8 if cfg is None else cfg.curve_steps
    return {OP_BRIGHTNESS: 1, OP_CONTRAST: 1, OP_SATURATION: 1, OP_COLOR: 3 * L, OP_INPAINT: 1,
            OP_TONE: L, OP_SHARPNESS: 1, OP_WHITE: 1, OP_EXPOSURE: 1, OP_WHITEBALANCE: 3,
            OP_BNW: 1, OP_BLUR: 1, OP_HUE: 1}[op_id]


# ---------------------------------------------------------------- pixel helpers
def lerp(a, b, l):
    """utils/operator_utils.py:5-6"""
    return (1 - l) * a + l * b


def rgb2lum(image):
    """utils/operator_utils.py:9-11 (weights 0.27 / 0.67 / 0.06)."""
    lum = 0.27 * image[:, 0, :, :] + 0.67 * image[:, 1, :, :] + 0.06 * image[:, 2, :, :]
    return lum[:, None, :, :]


def tanh_range(l, r, initial=None):
    """utils/operator_utils.py:21-34"""
    def act(x):
        bias = 0
        if initial is not None:
            y = 2 * (initial - l) / (r - l) - 1
            bias = 0.5 * math.log((1 + y) / (1 - y))
        return (torch.tanh(x + bias) * 0.5 + 0.5) * (r - l) + l
    return act


def _bc(param):
    """(B, n) -> (B, n, 1, 1) as ``param.unsqueeze(-1).unsqueeze(-1)`` in every process()."""
    return param.unsqueeze(-1).unsqueeze(-1)


# ---------------------------------------------------------------- process() per operator
def process_exposure(img, param, cfg):
    """models/operators.py:209-210"""
    return img * torch.exp(_bc(param) * np.log(2))


def process_contrast(img, param, cfg):
    """models/operators.py:240-245"""
    lum = torch.min(torch.max(rgb2lum(img), torch.tensor(0.0)), torch.tensor(1.0))
    contrast_lum = -torch.cos(np.pi * lum) * 0.5 + 0.5
    contrast_img = img / (lum + 1e-6) * contrast_lum
    return lerp(img, contrast_img, _bc(param))


def process_brightness(img, param, cfg):
    """models/operators.py:277-283"""
    hsv = _hsv.rgb_to_hsv(img)
    h, s, v = torch.chunk(hsv, chunks=3, dim=1)
    v_out = (v * (1 + _bc(param))).clamp(0, 1)
    return _hsv.hsv_to_rgb(torch.cat([h, s, v_out], dim=1))


def process_saturation(img, param, cfg):
    """models/operators.py:473-479"""
    hsv = _hsv.rgb_to_hsv(img)
    h, s, v = torch.chunk(hsv, chunks=3, dim=1)
    s_out = (s * (1 + _bc(param))).clamp(0, 1)
    return _hsv.hsv_to_rgb(torch.cat([h, s_out, v], dim=1))


_LAPLACE = torch.tensor([[[[0, -1, 0], [-1, 4, -1], [0, -1, 0]]]], dtype=torch.float)


def process_sharpness(img, param, cfg):
    """models/operators.py:351-358 (per-channel 3x3 Laplacian, zero padding 1)."""
    planes = [F.conv2d(c, _LAPLACE.to(img.device), padding=1) for c in img.split([1, 1, 1], 1)]
    return img + _bc(param) * torch.cat(planes, 1)


def process_white(img, param, cfg):
    """models/operators.py:510-512"""
    return torch.ones_like(img)


def process_whitebalance(img, param, cfg):
    """models/operators.py:548-549"""
    return img * _bc(param)


def _curve(img, curve, steps, in_place_scale):
    # shared body of models/operators.py:578-585 (tone) and :608-616 (color)
    curve_sum = curve.sum(2) + 1e-10
    total = torch.zeros_like(img)
    for i in range(steps):
        total = total + torch.clamp(img - 1.0 * i / steps, 0, 1.0 / steps) * curve[:, :, i, :, :]
    if in_place_scale:
        return total * (steps / curve_sum)         # color: total_img *= L / sum   (:615)
    return total * steps / curve_sum               # tone:  total_img * L / sum    (:584)


def process_tone(img, param, cfg):
    """models/operators.py:571-585 (one curve shared by the three channels)."""
    L = cfg.curve_steps
    return _curve(img, param.view(-1, 1, L, 1, 1), L, False)


def process_color(img, param, cfg):
    """models/operators.py:607-616 (one curve per channel, index c*L+i)."""
    L = cfg.curve_steps
    return _curve(img, param.view(-1, 3, L, 1, 1), L, True)


def process_bnw(img, param, cfg):
    """models/operators.py:314-316"""
    return lerp(img, rgb2lum(img), _bc(param))


def _gaussian_kernel():
    """models/operators.py:685-709 (kernel_size 3, sigma 2): the 3x3 weights of get_gaussian_kernel."""
    x_coord = torch.arange(3)
    x_grid = x_coord.repeat(3).view(3, 3)
    xy_grid = torch.stack([x_grid, x_grid.t()], dim=-1).float()
    mean, variance = (3 - 1) / 2., 2 ** 2.
    k = (1. / (2. * np.pi * variance)) * torch.exp(-torch.sum((xy_grid - mean) ** 2., dim=-1) / (2 * variance))
    return (k / torch.sum(k)).view(1, 1, 3, 3)


_GAUSS = _gaussian_kernel()


def process_blur(img, param, cfg):
    """models/operators.py:397-404 (per-channel 3x3 Gaussian, zero padding 1, then lerp)."""
    planes = [F.conv2d(c, _GAUSS.to(img.device), padding=1) for c in img.split([1, 1, 1], 1)]
    return lerp(img, torch.cat(planes, 1), _bc(param))


def process_hue(img, param, cfg):
    """models/operators.py:432-438, with the parameter broadcast per batch row (the reference's
    `param.expand_as(value)` of a (B, 1) parameter only works for B == 1)."""
    hsv = _hsv.rgb_to_hsv(img)
    hue, sat, value = hsv.split([1, 1, 1], 1)
    out_hsv = torch.cat((_bc(param).expand_as(value), sat, value), 1)
    return _hsv.hsv_to_rgb(out_hsv)


_PROCESS = {
    OP_BNW: process_bnw, OP_BLUR: process_blur, OP_HUE: process_hue,
    OP_BRIGHTNESS: process_brightness, OP_CONTRAST: process_contrast, OP_SATURATION: process_saturation,
    OP_COLOR: process_color, OP_TONE: process_tone, OP_SHARPNESS: process_sharpness, OP_WHITE: process_white,
    OP_EXPOSURE: process_exposure, OP_WHITEBALANCE: process_whitebalance,
}


def process(op_id, img, param, cfg=None):
    cfg = cfg or default_cfg()
    return _PROCESS[op_id](img, param, cfg)


def execute(op_id, img, param, mask=None, cfg=None):
    """Operator.execute with a specified parameter: models/operators.py:112-131, dispatched as
    executors/executor.py:33-55 does (op_ind < 0 is the identity, no clamp)."""
    if op_id < 0:
        return img
    if mask is None:
        mask = torch.ones_like(img)
    out = process(op_id, img, param, cfg)
    out = out * mask + img * (1 - mask)
    return torch.clamp(out, 0, 1)


def chain(img, op_ids, params, mask=None, cfg=None):
    """K successive Executor.execute calls, as the planner replays a sequence
    (utils/beam_search.py:165-167 applied along one beam)."""
    for op_id, p in zip(op_ids, params):
        img = execute(op_id, img, p, mask, cfg)
    return img


def l1_dist(x1, x2):
    """get_dist(..., 'L1'): utils/beam_search.py:170-173 -- one scalar over the whole batch."""
    return (x1 - x2).norm(1) / x1.numel()


def l1_mean(pred, target):
    """training / eval L1: experiments/t2onet/train_seq2seqL1.py:85, utils/eval.py:50-52."""
    return torch.abs(pred - target).mean()


# ---------------------------------------------------------------- parameter regressors
def regress(op_id, feat, cfg=None):
    """op_param_regressor of each class (applied to the fc2 output)."""
    cfg = cfg or default_cfg()
    if op_id == OP_EXPOSURE:        # :193-196
        return tanh_range(-cfg.exposure_range, cfg.exposure_range, initial=0)(feat)
    if op_id == OP_CONTRAST:        # :231-232
        return torch.tanh(feat)
    if op_id == OP_BRIGHTNESS:      # :266-269
        return tanh_range(-cfg.brightness_range, cfg.brightness_range, initial=0)(feat)
    if op_id == OP_SHARPNESS:       # :340-343
        return torch.sigmoid(feat) * cfg.sharpness_range
    if op_id == OP_SATURATION:      # :461-465
        return torch.tanh(F.relu(feat)) * cfg.saturation_range[1] + torch.tanh(F.relu(-feat)) * cfg.saturation_range[0]
    if op_id == OP_WHITE:           # :501-502
        return torch.sigmoid(feat)
    if op_id == OP_WHITEBALANCE:    # :535-546
        m = torch.tensor([0, 1, 1], dtype=torch.float).view(1, 3)
        scaling = torch.exp(tanh_range(-0.5, 0.5)(feat * m))
        return scaling * 1.0 / (1e-5 + 0.27 * scaling[:, 0] + 0.67 * scaling[:, 1] + 0.06 * scaling[:, 2]).unsqueeze(1)
    if op_id in (OP_TONE, OP_COLOR):  # :566-567, :602-603
        return feat
    raise KeyError(op_id)


def param_range(op_id, cfg=None):
    """get_param_range() -> (ub, lb, initial)  (note the order) of each class."""
    cfg = cfg or default_cfg()
    if op_id == OP_EXPOSURE:
        return cfg.exposure_range, -cfg.exposure_range, 0
    if op_id == OP_CONTRAST:
        return 1, -1, 0
    if op_id == OP_BRIGHTNESS:
        return cfg.brightness_range, -cfg.brightness_range, 0
    if op_id == OP_SHARPNESS:
        return cfg.sharpness_range, 0, cfg.sharpness_range / 2
    if op_id == OP_SATURATION:
        return cfg.saturation_range[1], cfg.saturation_range[0], 0
    if op_id == OP_WHITE:
        return 1, 0, 0.5
    if op_id == OP_WHITEBALANCE:
        return 1.8, 0.4, (0.4 + 1.8) / 2
    if op_id == OP_TONE:
        return cfg.tone_curve_range[1], cfg.tone_curve_range[0], (cfg.tone_curve_range[1] + cfg.tone_curve_range[0]) / 2
    if op_id == OP_COLOR:
        return cfg.color_curve_range[1], cfg.color_curve_range[0], (cfg.color_curve_range[1] + cfg.color_curve_range[0]) / 2
    if op_id == OP_INPAINT:
        return 0, 0, 0
    raise KeyError(op_id)


class OracleExecutor:
    """executors/executor.py:14-63 restated over ``execute`` above (specified_param path only)."""

    def __init__(self, cfg=None):
        self.cfg = cfg or default_cfg()

    def execute(self, img, op_ind, mask, features=None, specified_param=None, has_noise=False):
        if op_ind < 0:
            return img, torch.zeros(img.shape[0], 24, dtype=torch.float)
        assert specified_param is not None and features is None and not has_noise
        return execute(op_ind, img, specified_param, mask, self.cfg), specified_param

    def get_param_bnd(self, op_ind):
        return param_range(op_ind, self.cfg)

    def get_param_num(self, op_ind):
        return num_params(op_ind, self.cfg)
