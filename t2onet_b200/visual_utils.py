"""Device-side image <-> tensor conversions: utils/visual_utils.py of the reference (img2tensor :61-70, tensor2img
:50-58, the `/ 255` of load_train_img / load_infer_img :20-47) on the GPU, bit-identical to the host arithmetic
(IEEE x / 255; (x * 255) truncated to uint8).  Images cross PCIe as the uint8 arrays they are -- a quarter of the
float32 bytes -- and the conversion runs at HBM speed (t2o_u8_to_f32 / t2o_f32_to_u8 / t2o_img2tensor / t2o_tensor2img).
No CPU fallback: a CPU tensor raises T2OError."""
import numpy as np
import torch

from . import _lib


def _req(t, dtype, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.T2OError('%s: a CUDA tensor is required (no CPU fallback)' % what)
    if t.dtype != dtype:
        raise _lib.T2OError('%s: dtype %s expected, got %s' % (what, dtype, t.dtype))
    return t.contiguous()


def u8_to_float(x_u8, out=None):
    """uint8 CUDA tensor of any shape -> float32 tensor of the same shape, x / 255 (planar layout kept)."""
    x = _req(x_u8, torch.uint8, 'u8_to_float')
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    elif out.shape != x.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise _lib.T2OError('u8_to_float: `out` must be a contiguous float32 tensor of the input shape on the same device')
    _lib.check(_lib.lib().t2o_u8_to_f32(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(x.device)))
    return out


def float_to_u8(x, out=None):
    """float32 CUDA tensor -> uint8 tensor of the same shape: (x * 255) truncated, clamped to [0, 255]."""
    x = _req(x, torch.float32, 'float_to_u8')
    if out is None:
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    elif out.shape != x.shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != x.device:
        raise _lib.T2OError('float_to_u8: `out` must be a contiguous uint8 tensor of the input shape on the same device')
    _lib.check(_lib.lib().t2o_f32_to_u8(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(x.device)))
    return out


def img2tensor(img, device=None):
    """utils/visual_utils.py:61-70: BGR image (H, W, 3) uint8 -- numpy array, CPU or CUDA tensor; a batch (N, H, W, 3)
    is accepted too -- -> RGB tensor (1 | N, 3, H, W) float32 in [0, 1] on the GPU.  Only the uint8 bytes are copied
    to the device."""
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img))
    if not img.is_cuda:
        img = img.to(device if device is not None else 'cuda', non_blocking=True)
    img = _req(img, torch.uint8, 'img2tensor')
    if img.dim() == 3:
        img = img.unsqueeze(0)
    if img.dim() != 4 or img.shape[-1] != 3:
        raise _lib.T2OError('img2tensor: (H, W, 3) or (N, H, W, 3) uint8 expected, got %s' % (tuple(img.shape),))
    N, H, W, _ = img.shape
    out = torch.empty(N, 3, H, W, dtype=torch.float32, device=img.device)
    _lib.check(_lib.lib().t2o_img2tensor(_lib.ptr(img), _lib.ptr(out), N, H, W, _lib.stream_ptr(img.device)))
    return out


def tensor2img_device(tensor):
    """(N, 3, H, W) float32 RGB CUDA tensor -> (N, H, W, 3) uint8 BGR CUDA tensor (the device half of tensor2img)."""
    t = _req(tensor, torch.float32, 'tensor2img')
    if t.dim() != 4 or t.shape[1] != 3:
        raise _lib.T2OError('tensor2img: (N, 3, H, W) expected, got %s' % (tuple(t.shape),))
    N, _, H, W = t.shape
    out = torch.empty(N, H, W, 3, dtype=torch.uint8, device=t.device)
    _lib.check(_lib.lib().t2o_tensor2img(_lib.ptr(t), _lib.ptr(out), N, H, W, _lib.stream_ptr(t.device)))
    return out


def tensor2img(tensor):
    """utils/visual_utils.py:50-58: (1, 3, H, W) float RGB tensor -> (H, W, 3) uint8 BGR numpy image; only the uint8
    bytes travel back to the host."""
    return tensor2img_device(tensor)[0].cpu().numpy()
