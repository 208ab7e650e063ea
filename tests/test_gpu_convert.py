"""8-bit image <-> float tensor conversions on the device against utils/visual_utils.py's host arithmetic (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_img2tensor(img):
    """utils/visual_utils.py:61-70, verbatim semantics (host)."""
    img = img[:, :, ::-1]
    t = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))) / 255
    return t.unsqueeze(0)


def ref_tensor2img(tensor):
    """utils/visual_utils.py:50-58"""
    out = tensor.squeeze(0).permute(1, 2, 0) * 255
    return out.cpu().numpy().astype(np.uint8)[:, :, ::-1]


def test_all_256_values_divide_exactly():
    from t2onet_b200 import visual_utils as V
    u = torch.arange(256, dtype=torch.uint8)
    got = V.u8_to_float(u.cuda()).cpu()
    assert torch.equal(got, u / 255)                                       # torch true-divide on the host: the reference's x / 255
    assert not torch.equal(got, u.float() * np.float32(1 / 255))           # ... which a multiplication by 1/255 is not
    back = V.float_to_u8(got.cuda()).cpu()
    assert torch.equal(back, torch.from_numpy((got.numpy() * 255).astype(np.uint8)))


@pytest.mark.parametrize('shape', [(1, 1, 1), (7, 5, 3), (64, 96, 3), (33, 130, 3), (128, 128, 3), (250, 333, 3)])
def test_img2tensor_and_back_match_host(shape):
    from t2onet_b200 import visual_utils as V
    H, W, _ = shape
    rng = np.random.RandomState(H * 1000 + W)
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    t = V.img2tensor(img)
    assert t.is_cuda and tuple(t.shape) == (1, 3, H, W)
    assert torch.equal(t.cpu(), ref_img2tensor(img))
    # tensor2img of arbitrary floats in [0, 1]
    x = torch.from_numpy(rng.rand(1, 3, H, W).astype(np.float32))
    x.view(-1)[:4] = torch.tensor([0.0, 1.0, 0.5, 0.999999])[:min(4, x.numel())]
    got = V.tensor2img(x.cuda())
    assert got.dtype == np.uint8 and got.shape == (H, W, 3)
    assert np.array_equal(got, ref_tensor2img(x))
    # batches and the planar pair
    batch = rng.randint(0, 256, size=(3, H, W, 3)).astype(np.uint8)
    tb = V.img2tensor(batch)
    for b in range(3):
        assert torch.equal(tb[b:b + 1].cpu(), ref_img2tensor(batch[b]))
    planar = torch.from_numpy(rng.randint(0, 256, size=(2, 3, H, W)).astype(np.uint8))
    assert torch.equal(V.u8_to_float(planar.cuda()).cpu(), planar / 255)
    xb = torch.from_numpy(rng.rand(2, 3, H, W).astype(np.float32))
    assert torch.equal(V.float_to_u8(xb.cuda()).cpu(), torch.from_numpy((xb.numpy() * 255).astype(np.uint8)))


def test_full_resolution_roundtrip_properties():
    """16x3x2048x3072-sized planes (BASELINE config 4's shape, one image here): u8 -> float -> u8 is the identity and the
    float image is what the chain kernels consume."""
    from t2onet_b200 import visual_utils as V
    g = torch.Generator(device='cuda').manual_seed(3)
    u = torch.randint(0, 256, (2, 3, 2048, 3072), dtype=torch.uint8, device='cuda', generator=g)
    f = V.u8_to_float(u)
    assert float(f.min()) >= 0.0 and float(f.max()) <= 1.0
    assert torch.equal(V.float_to_u8(f), u)                                # uint8(float(v) / 255 * 255) == v for all 256 values
    hwc = V.tensor2img_device(f)
    assert torch.equal(V.img2tensor(hwc), f)


def test_no_cpu_fallback():
    from t2onet_b200 import visual_utils as V, T2OError
    with pytest.raises(T2OError):
        V.u8_to_float(torch.zeros(4, dtype=torch.uint8))
    with pytest.raises(T2OError):
        V.float_to_u8(torch.zeros(4))
