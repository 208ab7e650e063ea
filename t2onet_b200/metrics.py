"""Evaluation-side pixel metrics of the reference on the GPU: `ssim` (utils/ssim/__init__.py:63-73, the pytorch_ssim
recipe) and the L1 of utils/eval.py:50-52, each in one pass over HBM (t2o_ssim_sum / t2o_l1_sum)."""
import torch

from . import _lib
from . import functional as TF


def ssim_sum(img1, img2):
    """Per-image sums of the SSIM map -> (B,) float32 CUDA tensor; window 11, sigma 1.5, zero padding."""
    _lib.require_cuda(img1, img2)
    if img1.shape != img2.shape or img1.dim() != 4:
        raise _lib.T2OError('ssim: two (B, C, H, W) tensors of one shape expected, got %s and %s' % (tuple(img1.shape), tuple(img2.shape)))
    img1, img2 = img1.contiguous(), img2.contiguous()
    B, C, H, W = img1.shape
    lib = _lib.lib()
    out = torch.empty(B, device=img1.device, dtype=torch.float32)
    ws = _lib.workspace(img1.device, lib.t2o_ssim_workspace_bytes(B, C, H, W))
    _lib.check(lib.t2o_ssim_sum(_lib.ptr(img1), _lib.ptr(img2), _lib.ptr(out), B, C, H, W, _lib.ptr(ws), ws.numel(),
                                _lib.stream_ptr(img1.device)))
    return out


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/ssim/__init__.py:63-73: mean of the SSIM map over everything (size_average) or per image."""
    if window_size != 11:
        raise _lib.T2OError('ssim: only the reference\'s window_size = 11 is implemented')
    s = ssim_sum(img1, img2)
    per = float(img1[0].numel())
    return s.sum() / (per * img1.shape[0]) if size_average else s / per


class SSIM(torch.nn.Module):
    """utils/ssim/__init__.py:44-61"""

    def __init__(self, window_size=11, size_average=True):
        super(SSIM, self).__init__()
        self.window_size, self.size_average = window_size, size_average

    def forward(self, img1, img2):
        return ssim(img1, img2, self.window_size, self.size_average)


def l1(pred, gt):
    """utils/eval.py:50-52: torch.abs(pred - gt).mean() as a 0-dim tensor."""
    return TF.l1_sum(pred, gt).sum() / pred.numel()
