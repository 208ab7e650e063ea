"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import ops as O

TOL_PIX = 1e-5      # north star: max-abs on edited pixels and on the L1 (fp32)
TOL_GRAD = 1e-4     # north star: relative on parameter gradients


def sample_params(op, B, g, wide=False):
    n = O.num_params(op)
    u = torch.rand(B, n, generator=g)
    if op == O.OP_BRIGHTNESS:
        return (u * 0.6 - 0.3) if not wide else (u * 4 - 2)
    if op == O.OP_CONTRAST:
        return (u - 0.5) if not wide else (u * 2 - 1)
    if op == O.OP_SATURATION:
        return (u - 0.2) if not wide else (u * 3 - 1.5)
    if op == O.OP_COLOR:
        return 0.9 + 0.2 * u
    if op == O.OP_TONE:
        return 0.5 + 1.5 * u
    if op == O.OP_SHARPNESS:
        return u * 1.5
    if op == O.OP_EXPOSURE:
        return u * 2 - 1
    if op == O.OP_WHITEBALANCE:
        return 0.4 + 1.4 * u
    if op == O.OP_HUE:
        return (u * 6.2831853) if not wide else (u * 21 - 7)
    if op in (O.OP_BNW, O.OP_BLUR):
        return u if not wide else (u * 2 - 0.5)
    return u


def rel_err(a, b, atol=1e-7):
    """max|a-b| relative to max|b|; differences below `atol` (fp32 cancellation noise on gradients that are
    analytically ~0, e.g. contrast on an all-white region: R - 1 = -1e-6) count as zero."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.abs(a - b).max()
    return 0.0 if d <= atol else float(d / (np.abs(b).max() + 1e-12))


def rel_err_kinks(a, b, max_kinks=4, atol=1e-7):
    """rel_err for IMAGE gradients that ignores up to `max_kinks` elements: a pixel whose forward value lands within
    an ulp of a clamp edge / curve knot can fall on the other side of the kink in the kernels' closed forms than in
    the reference's op-by-op evaluation; its gradient then legitimately differs (both are one-sided derivatives)."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    d = np.sort(np.abs(a - b))
    d = d[:max(1, d.size - max_kinks)].max()
    return 0.0 if d <= atol else float(d / (np.abs(b).max() + 1e-12))


def kink_pixels(gimg, gimg_ref, rtol=1e-4):
    """Number of spatial positions (b, y, x) whose IMAGE gradient differs from the reference's by more than `rtol` of
    the gradient scale: pixels whose forward value sits within an ulp of a kink (clamp edge, curve knot, out == target)
    and falls on its other side in the kernels' closed forms.  Image gradients are per pixel (a kink before a stencil
    still only changes the derivative of that pixel's own operators), so this counts the kink pixels exactly."""
    a, b = np.asarray(gimg, dtype=np.float64), np.asarray(gimg_ref, dtype=np.float64)
    bad = (np.abs(a - b) > rtol * np.abs(b).max()).any(axis=1)
    return int(bad.sum())


def kink_slack(gimg, gimg_ref, numel, max_kinks=2, gain=2.0):
    """Absolute slack a mean-L1 PARAMETER gradient is allowed because of kink pixels: each one moves it by at most
    |dL1/dout| * |dout/dp| <= gain / numel.  Zero kink pixels -> no slack beyond fp32 cancellation noise."""
    n = kink_pixels(gimg, gimg_ref)
    assert n <= max_kinks, '%d pixels with mismatching image gradients (more than kinks explain)' % n
    return max(1e-7, n * gain / float(numel))


def max_abs(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max())


def oracle_chain_with_grads(img, ops, params, target, mask=None, wgt=None):
    """CPU oracle: out, per-image L1 sums, and gradients of loss w.r.t. params and img, where
    loss = (out*wgt).sum() if wgt is given else mean |out - target|."""
    x = img.clone().requires_grad_()
    ps = [p.clone().requires_grad_() for p in params]
    out = O.chain(x, ops, ps, mask)
    l1 = (out - target).abs().flatten(1).sum(1) if target is not None else None
    loss = (out * wgt).sum() if wgt is not None else O.l1_mean(out, target)
    loss.backward()
    gps = [p.grad if p.grad is not None else torch.zeros_like(p) for p in ps]
    return out.detach(), None if l1 is None else l1.detach(), gps, x.grad
