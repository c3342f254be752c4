"""Colour fix of the restored image (reference: SUPIR/utils/colorfix.py:45-120), on the GPU through the C ABI.

`wavelet_reconstruction(content, style)` keeps the high-frequency part of `content` (5 a-trous levels) and the low-frequency
part of `style`; `adaptive_instance_normalization(content, style)` matches per-channel mean / std. Inputs are fp32 NCHW
CUDA tensors as `SUPIRModel.batchify_sample` hands them over (SUPIR_model.py:131-135)."""
import torch

from . import _native
from ._native import call
from .ops import _need_cuda, _ptr, _stream


def wavelet_decomposition(image, levels=5):
    """colorfix.py:94-106: returns (high_freq, low_freq)."""
    _need_cuda(image)
    image = image.contiguous().float()
    N, C, H, W = image.shape
    high = torch.empty_like(image)
    cur, nxt = image, torch.empty_like(image)
    for i in range(levels):
        call("supir_wavelet_level", _ptr(cur), _ptr(nxt), _ptr(high), N * C, H, W, 2 ** i, int(i > 0), _stream())
        if cur is image:
            cur, nxt = nxt, torch.empty_like(image)
        else:
            cur, nxt = nxt, cur
    return high, cur


def wavelet_reconstruction(content_feat, style_feat):
    """colorfix.py:108-120."""
    content_high, _ = wavelet_decomposition(content_feat)
    _, style_low = wavelet_decomposition(style_feat)
    return _add(content_high, style_low)


def _add(a, b):
    from . import ops
    out = torch.empty_like(a)
    ops.axpby_f32(a.contiguous(), 1.0, b.contiguous(), 1.0, out)
    return out


def _plane_stats(x):
    N, C = x.shape[:2]
    planes = N * C
    hw = x.numel() // planes
    n = int(_native.load().supir_plane_stats_workspace(planes))
    ws = torch.empty(n, dtype=torch.float64, device=x.device)
    call("supir_plane_stats", _ptr(x), planes, hw, _ptr(ws), n, _stream())
    return ws, planes, hw


def adaptive_instance_normalization(content_feat, style_feat):
    """colorfix.py:59-71."""
    _need_cuda(content_feat, style_feat)
    content = content_feat.contiguous().float()
    style = style_feat.contiguous().float()
    assert content.dim() == 4 and content.shape[:2] == style.shape[:2], "The input feature should be 4D tensor."
    cs, planes, hw = _plane_stats(content)
    ss, _, _ = _plane_stats(style)
    out = torch.empty_like(content)
    call("supir_adain_apply", _ptr(content), _ptr(cs), _ptr(ss), _ptr(out), planes, hw, _stream())
    return out
