"""Record SSIM values of the UNMODIFIED reference (utils/ssim/__init__.py) -- TEST INFRASTRUCTURE ONLY.
    python -m oracle.make_ssim_golden        # writes tests/golden/ssim.npz
Also asserts that oracle/metrics.py reproduces them bit for bit."""
import os
import sys

import numpy as np
import torch

from . import metrics as OM
from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    if not ref_shims.available():
        sys.exit('reference tree not present')
    sys.path.insert(0, '/root/reference')
    from utils import ssim as RS
    g = torch.Generator().manual_seed(10 + 5000)
    rec = {}
    for name, (B, C, H, W) in {'a': (2, 3, 37, 70), 'b': (1, 3, 64, 128), 'c': (3, 1, 20, 9)}.items():
        x = torch.rand(B, C, H, W, generator=g)
        y = (x + 0.2 * torch.randn(B, C, H, W, generator=g)).clamp(0, 1)
        if name == 'b':
            y = x.clone()
        r_all, r_per = RS.ssim(x, y).item(), RS.ssim(x, y, size_average=False).numpy()
        assert r_all == OM.ssim(x, y).item() and np.array_equal(r_per, OM.ssim(x, y, size_average=False).numpy()), name
        rec[name + '_x'], rec[name + '_y'] = x.numpy(), y.numpy()
        rec[name + '_mean'], rec[name + '_per'] = np.float32(r_all), r_per
    np.savez_compressed(os.path.join(OUT, 'ssim.npz'), **rec)
    print('wrote ssim.npz')


if __name__ == '__main__':
    main()
