"""Drop-in `Operator` classes: same names, constructor (cfg), attributes and methods as
/root/reference/models/operators.py, with `execute` / `process` running the sm_100a kernels.

What stays PyTorch (as in the reference): the parameter heads -- fc1 -> LeakyReLU -> fc2 ->
op_param_regressor (models/operators.py:43-55,73-88): tiny GEMMs that define the parameter domain
and carry the checkpoint keys `<name>_op.fc1/fc2.*`.  What is replaced: process + mask blend +
clamp (models/operators.py:112-131) and autograd through them -> one fused forward kernel and one
recompute-backward kernel per call (t2onet_b200/csrc/t2o_chain.cu).

Out of scope: the discrete-parameter path (`cfg.discrete_param`, unused default)
and InpaintOperator's EdgeConnect network (a stub keeps the module/parameter names).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as TF
from ._lib import T2OError


# utils/operator_utils.py:5-34 -- kept for callers that import them from the operator module
def lerp(a, b, l):
    return (1 - l) * a + l * b


def rgb2lum(image):
    image = 0.27 * image[:, 0, :, :] + 0.67 * image[:, 1, :, :] + 0.06 * image[:, 2, :, :]
    return image[:, None, :, :]


def tanh_range(l, r, initial=None):
    def activation(x):
        bias = 0
        if initial is not None:
            y = 2 * (initial - l) / (r - l) - 1
            bias = 0.5 * math.log((1 + y) / (1 - y))
        return (torch.tanh(x + bias) * 0.5 + 0.5) * (r - l) + l
    return activation


class Operator(nn.Module):
    """models/operators.py:26-183.  img passed in must be RGB in [0, 1], (bs, 3, h, w), float32, CUDA."""

    op_id = None   # kernel operator id (= Executor index where one exists)
    SHORT_NAME, N_PARAMS = None, None

    def __init__(self, cfg):
        super(Operator, self).__init__()
        self.cfg = cfg
        self.is_discrete = getattr(cfg, 'discrete_param', 0)
        if self.is_discrete:
            raise NotImplementedError('discrete operator parameters (cfg.discrete_param) are out of scope; '
                                      'the reference default is 0 (options/fiveK_base_options.py:41)')
        self.channels = 2 * cfg.hidden_size
        self.op_param = None
        self.param_sample_flag = False
        self.curve_steps = getattr(cfg, 'curve_steps', 8)
        # subclasses declare SHORT_NAME and N_PARAMS (an int, or a function of cfg); the FC head is built right away
        self.short_name = self.SHORT_NAME
        self.num_op_param = self.N_PARAMS(cfg) if callable(self.N_PARAMS) else self.N_PARAMS
        if self.num_op_param is not None:
            self.setup()

    def setup(self):
        """must be called by child class (models/operators.py:43-55)"""
        output_dim = self.get_num_op_param()
        self.fc1 = nn.Linear(self.channels, self.cfg.operator_fc_dim)
        self.lrelu = nn.LeakyReLU(inplace=True)
        self.fc2 = nn.Linear(self.cfg.operator_fc_dim, output_dim)
        self.dist, self.ub, self.lb, self.initial = self.get_param_noise_distribution()

    def get_param_noise(self, bs):
        noise = self.dist.sample([bs])
        noise = (F.relu(noise) * (self.ub - self.initial) + F.relu(-noise) * (self.initial - self.lb)) / 3 \
            * self.cfg.param_noise_factor
        return noise

    def set_param_sample_flag(self, flag):
        self.param_sample_flag = flag

    def get_short_name(self):
        assert self.short_name
        return self.short_name

    def get_num_op_param(self):
        assert self.num_op_param is not None, 'Must specify the number of parameter'
        return self.num_op_param

    def extract_parameters(self, features):
        features = self.fc1(features)
        features = self.lrelu(features)
        features = self.fc2(features)
        return self.op_param_regressor(features)

    def op_param_regressor(self, features):
        raise NotImplementedError

    def process(self, img, param):
        """process the whole image, no mask blend, no clamp (models/operators.py:108,128); forward only."""
        return TF.process_raw(img, self.op_id, param, self.curve_steps)

    def execute(self, img, mask=None, features=None, specified_param=None, has_noise=False):
        """models/operators.py:112-131 -- fused process + mask blend + clamp; sets self.param / self.mask."""
        assert (features is None) ^ (specified_param is None)
        if features is not None:
            param = self.extract_parameters(features)
        else:
            param = specified_param
        if has_noise:
            param_noise = self.get_param_noise(img.shape[0]).to(img.device)
            param = param + param_noise
            param = torch.clamp(param, self.lb, self.ub)
        self.param = param
        self._mask, self._mask_like = mask, (img if mask is None else None)
        if param.device != img.device:
            param = param.to(img.device)
        return TF.chain(img, [self.op_id], [param.float()], mask, self.curve_steps)

    @property
    def mask(self):
        """The mask of the last execute (models/operators.py:123-125 stores ones_like(img) when none was given; the
        kernels need no such tensor, so it is only materialised if somebody reads the attribute)."""
        m = getattr(self, '_mask', None)
        if m is None and getattr(self, '_mask_like', None) is not None:
            m = torch.ones_like(self._mask_like)
        return m

    @mask.setter
    def mask(self, value):
        self._mask, self._mask_like = value, None

    def param_loss_fn(self):
        return F.mse_loss

    def get_param(self):
        return self.op_param

    def visualize_op(self, img=None):
        pass

    def get_param_range(self):
        raise NotImplementedError

    def get_param_noise_distribution(self):
        ub, lb, initial = self.get_param_range()
        dist = torch.distributions.normal.Normal(torch.zeros(self.num_op_param), torch.ones(self.num_op_param))
        return dist, ub, lb, initial


class ExposureOperator(Operator):
    """models/operators.py:186-222"""
    op_id = TF.OP_EXPOSURE

    SHORT_NAME, N_PARAMS = 'exposure', 1

    def op_param_regressor(self, features):
        bnd = self.cfg.exposure_range
        return tanh_range(-bnd, bnd, initial=0)(features)

    def get_param_range(self):
        return self.cfg.exposure_range, -self.cfg.exposure_range, 0


class ContrastOperator(Operator):
    """models/operators.py:224-257"""
    op_id = TF.OP_CONTRAST

    SHORT_NAME, N_PARAMS = 'contrast', 1

    def op_param_regressor(self, features):
        return torch.tanh(features)

    def get_param_range(self):
        return 1, -1, 0


class BrightnessOperator(Operator):
    """models/operators.py:259-295"""
    op_id = TF.OP_BRIGHTNESS

    SHORT_NAME, N_PARAMS = 'brightness', 1

    def op_param_regressor(self, features):
        bnd = self.cfg.brightness_range
        return tanh_range(-bnd, bnd, initial=0)(features)

    def get_param_range(self):
        return self.cfg.brightness_range, -self.cfg.brightness_range, 0


class SharpnessOperator(Operator):
    """models/operators.py:332-370"""
    op_id = TF.OP_SHARPNESS

    SHORT_NAME, N_PARAMS = 'sharpness', 1
    kernel = torch.tensor([[[[0, -1, 0], [-1, 4, -1], [0, -1, 0]]]], dtype=torch.float)      # (models/operators.py:338; the kernels' stencil)

    def op_param_regressor(self, features):
        return torch.sigmoid(features) * self.cfg.sharpness_range

    def get_param_range(self):
        ub = self.cfg.sharpness_range
        return ub, 0, ub / 2


class SaturationOperator(Operator):
    """models/operators.py:454-491"""
    op_id = TF.OP_SATURATION

    SHORT_NAME, N_PARAMS = 'saturation', 1

    def op_param_regressor(self, features):
        return torch.tanh(F.relu(features)) * self.cfg.saturation_range[1] + \
            torch.tanh(F.relu(-features)) * self.cfg.saturation_range[0]

    def get_param_range(self):
        return self.cfg.saturation_range[1], self.cfg.saturation_range[0], 0


class WhiteOperator(Operator):
    """models/operators.py:494-524"""
    op_id = TF.OP_WHITE

    SHORT_NAME, N_PARAMS = 'color_bg', 1

    def op_param_regressor(self, features):
        return torch.sigmoid(features)

    def get_param_range(self):
        return 1, 0, 0.5


class ImprovedWhiteBalanceOperator(Operator):
    """models/operators.py:527-555"""
    op_id = TF.OP_WHITEBALANCE

    SHORT_NAME, N_PARAMS = 'whitebalance', 3

    def op_param_regressor(self, features):
        log_wb_range = 0.5
        mask = torch.tensor([0, 1, 1], dtype=torch.float).view(1, 3).to(features.device)
        features = features * mask
        color_scaling = torch.exp(tanh_range(-log_wb_range, log_wb_range)(features))
        color_scaling = color_scaling * 1.0 / (1e-5 + 0.27 * color_scaling[:, 0] + 0.67 * color_scaling[:, 1] +
                                               0.06 * color_scaling[:, 2]).unsqueeze(1)
        return color_scaling

    def get_param_range(self):
        return 1.8, 0.4, (0.4 + 1.8) / 2


class ToneOperator(Operator):
    """models/operators.py:557-591"""
    op_id = TF.OP_TONE

    SHORT_NAME = 'tone'

    @staticmethod
    def N_PARAMS(cfg):
        return cfg.curve_steps

    def op_param_regressor(self, features):
        return features

    def get_param_range(self):
        ub, lb = self.cfg.tone_curve_range[1], self.cfg.tone_curve_range[0]
        return ub, lb, (ub + lb) / 2


class ColorOperator(Operator):
    """models/operators.py:593-622"""
    op_id = TF.OP_COLOR

    SHORT_NAME = 'hue'

    @staticmethod
    def N_PARAMS(cfg):
        return 3 * cfg.curve_steps

    def op_param_regressor(self, features):
        return features

    def get_param_range(self):
        ub, lb = self.cfg.color_curve_range[1], self.cfg.color_curve_range[0]
        return ub, lb, (ub + lb) / 2


class BNWOperator(Operator):
    """models/operators.py:298-329 (no Executor slot): lerp(img, luminance, p)"""
    op_id = TF.OP_BNW

    SHORT_NAME, N_PARAMS = 'black&white', 1

    def op_param_regressor(self, features):
        return torch.sigmoid(features)

    def get_param_range(self):
        return 1, 0, 0.5


class BlurOperator(Operator):
    """models/operators.py:373-411 (no Executor slot): lerp(img, 3x3 Gaussian(sigma 2, zero padding) * img, p)"""
    op_id = TF.OP_BLUR

    SHORT_NAME, N_PARAMS = 'blur', 1

    def op_param_regressor(self, features):
        return torch.sigmoid(features)

    def get_param_range(self):
        return 1, 0, 0.5


class HueOperator(Operator):
    """models/operators.py:414-451 (no Executor slot): hsv_to_rgb(param, s, v) -- the hue of every pixel replaced by the
    parameter (radians, kornia's [0, 2 pi) scale).  The reference's `param.expand_as(value)` only broadcasts for a
    batch of one; here every batch row takes its own parameter."""
    op_id = TF.OP_HUE

    SHORT_NAME, N_PARAMS = 'hue_', 1

    def op_param_regressor(self, features):
        return features

    def get_param_range(self):
        return 1, 0, 0.5


class InpaintOperator(Operator):
    """models/operators.py:625-682.  The EdgeConnect inpainting network is OUT OF SCOPE (a local,
    deep-CNN operator).  The stub keeps the module and checkpoint names."""
    op_id = TF.OP_INPAINT

    SHORT_NAME, N_PARAMS = 'inpaint_obj', 1

    def op_param_regressor(self, features):
        return torch.zeros((features.shape[0], self.num_op_param), requires_grad=True, device=features.device)

    def param_loss_fn(self):
        def psudo_loss_fn(pred, tgt):
            return 0
        return psudo_loss_fn

    def get_param_range(self):
        return 0, 0, 0

    def process(self, img, param):
        raise NotImplementedError('InpaintOperator (EdgeConnect) is outside the B200 hot path')

    def execute(self, img, mask=None, features=None, specified_param=None, has_noise=False):
        raise NotImplementedError('InpaintOperator (EdgeConnect) is outside the B200 hot path')
