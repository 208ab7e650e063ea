"""RGB <-> HSV as kornia (<= 0.5.x) computes it -- TEST INFRASTRUCTURE ONLY.

The reference calls ``kornia.rgb_to_hsv`` / ``kornia.hsv_to_rgb``
(/root/reference/models/operators.py:278,282 Brightness; :474,478 Saturation;
:432,438 Hue).  kornia is a third-party dependency that is neither vendored nor
pinned (``requirements.txt:5``; the top-level ``kornia.rgb_to_hsv`` export only
exists up to kornia 0.5.x) and is not installed in this image, so this module
restates the published algorithm of ``kornia/color/hsv.py`` (0.4.1 - 0.5.x):

rgb_to_hsv(image, eps=1e-6)
    v      = max_c(image)                      (first max index on ties)
    delta  = v - min_c(image)
    s      = delta / (v + eps)
    delta' = where(delta == 0, 1, delta)
    h      = select_by_argmax([bc-gc, 2*delta'+rc-bc, 4*delta'+gc-rc]) / delta'
             with xc = v - x
    h      = ((h / 6) mod 1) * 2*pi

hsv_to_rgb(image)
    h  = h / (2*pi);  hi = floor(6h) mod 6;  f = (6h mod 6) - hi
    p  = v(1-s);  q = v(1-f s);  t = v(1-(1-f)s)
    rgb = [(v,q,p,p,t,v), (t,v,v,q,p,p), (p,p,t,v,v,q)][hi]

Parity status: **unpinned** (no kornia here to run against; older kornia 0.2-0.4.0
used ``s = delta / v`` with a NaN/1e-31 guard, which moves Brightness by <= 1.3e-6,
SURVEY.md section 8c).
"""
import math

import torch


def rgb_to_hsv(image: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """(*, 3, H, W) RGB in [0, 1] -> (*, 3, H, W) with h in [0, 2pi), s, v."""
    maxc, _ = image.max(-3)
    is_max = image == maxc.unsqueeze(-3)
    # index of the first channel that attains the maximum
    _, max_idx = ((is_max.cumsum(-3) == 1) & is_max).max(-3)
    minc = image.min(-3)[0]

    v = maxc
    delta = maxc - minc
    s = delta / (v + eps)

    # avoid a division by zero for gray pixels (the hue is irrelevant there)
    delta = torch.where(delta == 0, torch.ones_like(delta), delta)

    dist = maxc.unsqueeze(-3) - image
    rc = dist[..., 0, :, :]
    gc = dist[..., 1, :, :]
    bc = dist[..., 2, :, :]

    h = torch.stack([bc - gc, 2.0 * delta + rc - bc, 4.0 * delta + gc - rc], dim=-3)
    h = torch.gather(h, dim=-3, index=max_idx[..., None, :, :]).squeeze(-3)
    h = h / delta
    h = (h / 6.0) % 1.0
    h = 2 * math.pi * h
    return torch.stack([h, s, v], dim=-3)


def hsv_to_rgb(image: torch.Tensor) -> torch.Tensor:
    """(*, 3, H, W) with h in [0, 2pi] -> RGB."""
    h = image[..., 0, :, :] / (2 * math.pi)
    s = image[..., 1, :, :]
    v = image[..., 2, :, :]

    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    one = torch.tensor(1.0, device=image.device, dtype=image.dtype)
    p = v * (one - s)
    q = v * (one - f * s)
    t = v * (one - (one - f) * s)

    hi = hi.long()
    idx = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    table = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(table, -3, idx)
