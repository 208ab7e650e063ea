"""Evaluation metrics on the GPU (t2o_ssim_sum / t2o_l1_sum through t2onet_b200.metrics) against the values recorded
from the reference's utils/ssim and against the oracle on ragged and larger shapes.  Tolerance: 1e-5 absolute on the
SSIM mean (the kernel filters separably in fp32; the reference sums the 121 taps of the 2-D window)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as OM

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope='module')
def M():
    from t2onet_b200 import metrics
    return metrics


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_ssim_matches_reference_golden(M, golden_dir, name):
    d = np.load(os.path.join(golden_dir, 'ssim.npz'))
    x, y = torch.from_numpy(d[name + '_x']).cuda(), torch.from_numpy(d[name + '_y']).cuda()
    assert abs(M.ssim(x, y).item() - float(d[name + '_mean'])) <= TOL
    per = M.ssim(x, y, size_average=False).cpu().numpy()
    assert np.abs(per - d[name + '_per']).max() <= TOL
    assert abs(M.SSIM()(x, y).item() - float(d[name + '_mean'])) <= TOL


@pytest.mark.parametrize('shape', [(1, 3, 1, 1), (2, 3, 33, 65), (1, 3, 600, 901), (4, 3, 128, 128), (1, 1, 5, 300)])
def test_ssim_matches_oracle(M, shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(*shape, generator=g)
    y = (x + 0.1 * torch.randn(*shape, generator=g)).clamp(0, 1)
    ref = OM.ssim(x, y, size_average=False)
    got = M.ssim(x.cuda(), y.cuda(), size_average=False).cpu()
    assert (got - ref).abs().max().item() <= TOL
    assert abs(M.ssim(x.cuda(), x.cuda()).item() - 1.0) <= TOL                  # identical images
    assert abs(M.l1(x.cuda(), y.cuda()).item() - (x - y).abs().mean().item()) <= 1e-6


def test_ssim_rejects_cpu_tensors(M):
    import t2onet_b200 as T
    with pytest.raises(T.T2OError):
        M.ssim(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
