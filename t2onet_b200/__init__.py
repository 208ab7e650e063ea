"""t2onet_b200 -- B200-native (sm_100a) implementation of T2ONet's data-parallel hot path:
the global image-editing operators, their executor, the fused operator chain with L1 and its
backward, and the operation planner's candidate scoring.  See DESIGN.md.

Public surface (mirrors the reference's Python API):
    t2onet_b200.operators   Operator subclasses            (models/operators.py)
    t2onet_b200.executor    Executor                       (executors/executor.py)
    t2onet_b200.planner     beam_search, get_dist, ...     (utils/beam_search*.py)
    t2onet_b200.plans       planner records on disk + reader (preprocess/gen_greedy_seqs_FiveK.py, datasets/FiveKdataset.py)
    t2onet_b200.visual_utils  img2tensor / tensor2img on the device (utils/visual_utils.py)
    t2onet_b200.functional  chain / chain_l1 / score_candidates over the C-ABI (include/t2o.h)
"""
from . import functional  # noqa: F401
from .executor import Executor  # noqa: F401
from .operators import (Operator, ExposureOperator, ContrastOperator, BrightnessOperator, SharpnessOperator,  # noqa: F401
                        SaturationOperator, WhiteOperator, ImprovedWhiteBalanceOperator, ToneOperator, ColorOperator,
                        InpaintOperator, BNWOperator, BlurOperator, HueOperator)
from . import planner  # noqa: F401
from . import plans  # noqa: F401
from . import visual_utils  # noqa: F401
from ._lib import T2OError, lib  # noqa: F401

__version__ = '0.1.0'


def default_options(**over):
    """Operator/executor options with the reference defaults (options/fiveK_base_options.py:30-54)."""
    from types import SimpleNamespace
    opt = SimpleNamespace(
        hidden_size=256, operator_fc_dim=512, discrete_param=0, discrete_step=10,
        exposure_range=3.5, sharpness_range=1.5, brightness_range=2, curve_steps=8,
        tone_curve_range=(0.5, 2), color_curve_range=(0.90, 1.10), saturation_range=(-0.2, 0.8),
        param_noise_factor=0.0, explore_prob=0.0)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt
