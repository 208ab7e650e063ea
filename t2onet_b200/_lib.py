"""ctypes binding of the C-ABI in include/t2o.h (t2onet_b200/lib/libt2o_b200.so).

There is no CPU or PyTorch-eager fallback: if the library cannot be loaded every operator call
raises.  Workspaces are cached per (device, stream) and zero-initialised once (the library leaves
them zeroed after every call).
"""
import ctypes
import os

import torch

from . import build as _build

MAX_CHAIN = 8
MAX_OP_PARAMS = 24
FLAG_RAW_PROCESS = 1

_STATUS = {1: 'invalid argument', 2: 'unsupported configuration', 3: 'workspace missing or too small',
           4: 'CUDA runtime error', 5: 'no usable device / driver entry point'}

_lib = None
_workspaces = {}

c_float_p = ctypes.c_void_p
c_int_p = ctypes.POINTER(ctypes.c_int)


class T2OError(RuntimeError):
    pass


class NMState(ctypes.Structure):
    """t2o_nm_state of include/t2o.h: device pointers of the Nelder-Mead state arrays."""
    _fields_ = [(n, ctypes.c_void_p) for n in ('sim', 'fsim', 'vec', 'fxr', 'xbest', 'fbest', 'perm', 'ctl')]


def _declare(lib):
    lib.t2o_version.restype = ctypes.c_int
    lib.t2o_status_string.restype = ctypes.c_char_p
    lib.t2o_status_string.argtypes = [ctypes.c_int]
    lib.t2o_last_cuda_error.restype = ctypes.c_char_p
    lib.t2o_num_params.restype = ctypes.c_int
    lib.t2o_num_params.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.t2o_workspace_bytes.restype = ctypes.c_size_t
    lib.t2o_workspace_bytes.argtypes = [ctypes.c_int] * 4
    lib.t2o_score_workspace_bytes.restype = ctypes.c_size_t
    lib.t2o_score_workspace_bytes.argtypes = [ctypes.c_int] * 4
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.t2o_chain_forward.restype = ci
    lib.t2o_chain_forward.argtypes = [ci, c_int_p, c_int_p, vp, vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci,
                                      vp, ctypes.c_size_t, vp]
    lib.t2o_chain_backward.restype = ci
    lib.t2o_chain_backward.argtypes = [ci, c_int_p, c_int_p, vp, vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp,
                                       ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.t2o_rows_forward.restype = ci
    lib.t2o_rows_forward.argtypes = [ci, vp, c_int_p, ci, vp, vp, ci, vp, ci, vp, vp, vp, vp, ci, ci, ci, ci,
                                     vp, ctypes.c_size_t, vp]
    lib.t2o_rows_backward.restype = ci
    lib.t2o_rows_backward.argtypes = [ci, vp, c_int_p, ci, vp, vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp,
                                      ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.t2o_l1_sum.restype = ci
    lib.t2o_l1_sum.argtypes = [vp, vp, vp, ci, ctypes.c_int64, vp, ctypes.c_size_t, vp]
    lib.t2o_score_candidates.restype = ci
    lib.t2o_score_candidates.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, ci, vp, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.t2o_score_candidates_masked.restype = ci
    lib.t2o_score_candidates_masked.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.t2o_ssim_workspace_bytes.restype = ctypes.c_size_t
    lib.t2o_ssim_workspace_bytes.argtypes = [ci] * 4
    lib.t2o_ssim_sum.restype = ci
    lib.t2o_ssim_sum.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
    for name in ('t2o_u8_to_f32', 't2o_f32_to_u8'):
        getattr(lib, name).restype = ci
        getattr(lib, name).argtypes = [vp, vp, ctypes.c_int64, vp]
    for name in ('t2o_img2tensor', 't2o_tensor2img'):
        getattr(lib, name).restype = ci
        getattr(lib, name).argtypes = [vp, vp, ci, ci, ci, vp]
    lib.t2o_nm_run_resident.restype = ci
    lib.t2o_nm_run_resident.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, ci, ci, ctypes.POINTER(NMState), ci, ctypes.c_float, vp, vp,
                                        vp, vp, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.t2o_topk_min.restype = ci
    lib.t2o_topk_min.argtypes = [vp, vp, ci, ci, vp, vp, vp]
    lib.t2o_nm_start.restype = ci
    lib.t2o_nm_start.argtypes = [ctypes.POINTER(NMState), ci, vp, vp, vp, vp, vp, vp]
    lib.t2o_nm_advance.restype = ci
    lib.t2o_nm_advance.argtypes = [ctypes.POINTER(NMState), ci, vp, ctypes.c_float, vp, vp, vp]
    return lib


EXPORTS = ['t2o_version', 't2o_status_string', 't2o_last_cuda_error', 't2o_num_params', 't2o_workspace_bytes',
           't2o_score_workspace_bytes', 't2o_chain_forward', 't2o_chain_backward', 't2o_rows_forward',
           't2o_rows_backward', 't2o_l1_sum',
           't2o_score_candidates', 't2o_score_candidates_masked', 't2o_topk_min', 't2o_nm_run_resident', 't2o_nm_start', 't2o_nm_advance', 't2o_ssim_workspace_bytes', 't2o_ssim_sum',
           't2o_u8_to_f32', 't2o_f32_to_u8', 't2o_img2tensor', 't2o_tensor2img']


def lib():
    """Load (building first if the sources are newer and nvcc is present) the native library."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if _build.is_stale():
            if _build.nvcc_path() is None and not os.path.exists(path):
                raise T2OError('native library %s is missing and nvcc is not available; '
                               'run `python -m t2onet_b200.build` where nvcc exists' % path)
            if _build.nvcc_path() is not None:
                _build.build()
        _lib = _declare(ctypes.CDLL(path))
    return _lib


def check(status):
    if status != 0:
        msg = _STATUS.get(status, 'status %d' % status)
        if status == 4:
            msg += ': ' + lib().t2o_last_cuda_error().decode()
        raise T2OError('t2o: ' + msg)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


def stream_ptr(device):
    """The stream handle every C entry point takes.  The kernels launch on the process's CURRENT device, so it is made the
    tensors' device first (a no-op under the one-process-per-GPU convention; without it a tensor on cuda:1 in a process whose
    current device is cuda:0 would fail with an invalid resource handle)."""
    device = torch.device(device)
    if device.index is not None and torch.cuda.current_device() != device.index:
        torch.cuda.set_device(device)
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(device, nbytes):
    """Zero-initialised scratch for the current stream of `device`, grown geometrically."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        size = max(int(nbytes), 1 << 20)
        if ws is not None:
            size = max(size, 2 * ws.numel())
        ws = torch.zeros(size, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_status_words = {}


def status_word(device):
    """Per-device uint32 the per-row kernels flag invalid rows in (bit 0); see functional.rows_status()."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    w = _status_words.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=device)
        _status_words[key] = w
    return w


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise T2OError('t2onet_b200 operators run on CUDA tensors only (no CPU fallback); got a %s tensor' % t.device)
        if t.dtype != torch.float32:
            raise T2OError('t2onet_b200 operators are float32; got %s' % t.dtype)
