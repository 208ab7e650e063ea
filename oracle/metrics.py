"""SSIM oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (torch fp32) of the SSIM the reference's evaluation uses: utils/ssim/__init__.py:8-41, called as
`ssim(img1, img2)` from utils/eval.py:57-60.  Same arithmetic in the same order, so the values are bit-identical to the
reference's (asserted when tests/golden/ssim.npz is recorded, oracle/make_ssim_golden.py):

  * window: w1[x] = exp(-(x - 5)^2 / (2 * 1.5^2)) for x = 0..10 evaluated in Python floats, stored as float32 and
    normalised by its float32 sum (:8-10); the 2-D window is the float32 outer product w1 w1^T, one copy per channel
    (:13-17);
  * five depthwise convolutions with zero padding 5 -- of a, b, a*a, b*b, a*b (:20-29);
  * map = (2 mu_a mu_b + C1)(2 cov + C2) / ((mu_a^2 + mu_b^2 + C1)(var_a + var_b + C2)), C1 = 0.01^2, C2 = 0.03^2 (:31-34);
  * mean over everything, or per image (:36-39).
"""
import math

import torch
import torch.nn.functional as F

WINDOW, SIGMA = 11, 1.5
C1, C2 = 0.01 ** 2, 0.03 ** 2


def window_2d(channels, like):
    taps = [math.exp(-(x - WINDOW // 2) ** 2 / float(2 * SIGMA ** 2)) for x in range(WINDOW)]
    w1 = torch.Tensor(taps)
    w1 = w1 / w1.sum()
    w2 = torch.outer(w1, w1).float()                      # == w1[:, None].mm(w1[None, :]): one product per element
    return w2.expand(channels, 1, WINDOW, WINDOW).contiguous().type_as(like)


def _blur(x, w):
    return F.conv2d(x, w, padding=WINDOW // 2, groups=x.size(1))


def ssim(a, b, window_size=WINDOW, size_average=True):
    assert window_size == WINDOW
    w = window_2d(a.size(1), a)
    mu_a, mu_b = _blur(a, w), _blur(b, w)
    mu_aa, mu_bb, mu_ab = mu_a.pow(2), mu_b.pow(2), mu_a * mu_b
    var_a = _blur(a * a, w) - mu_aa
    var_b = _blur(b * b, w) - mu_bb
    cov = _blur(a * b, w) - mu_ab
    ssim_map = ((2 * mu_ab + C1) * (2 * cov + C2)) / ((mu_aa + mu_bb + C1) * (var_a + var_b + C2))
    if size_average:
        return ssim_map.mean()
    return ssim_map.mean(1).mean(1).mean(1)
