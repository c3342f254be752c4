"""Image pre/post-processing around the sampler (reference: SUPIR/util.py:60-94).

`PIL2Tensor` is host-side input preparation (a PIL resize) and is kept as the reference has it. `Tensor2PIL` — bicubic
resize back to the requested size, scale to 0..255, clip, uint8, HWC — sits right after the final VAE decode and runs as one
CUDA kernel (supir_image_to_uint8_bicubic); only the uint8 image crosses PCIe."""
import numpy as np
import torch

from ._native import call
from .ops import _need_cuda, _ptr, _stream


def _target_size(size, upscale, min_size, fix_resize):
    """Working size (multiples of 64) and requested output size of PIL2Tensor, with the reference's float operation order
    (SUPIR/util.py:65-79): scale, report the rounded size, grow the short side to `min_size`, optionally pin the short side to
    `fix_resize` (which also redefines the reported size), snap to 64."""
    dims = [float(size[0]) * upscale, float(size[1]) * upscale]          # (w, h)
    report = tuple(round(d) for d in dims)
    short = min(dims)
    if short < min_size:
        grow = min_size / short
        dims = [d * grow for d in dims]
    if fix_resize is not None:
        pin = fix_resize / min(dims)
        dims = [d * pin for d in dims]
        report = tuple(round(d) for d in dims)
    work = tuple(int(np.round(d / 64.0)) * 64 for d in dims)
    return work, report


def PIL2Tensor(img, upsacle=1, min_size=1024, fix_resize=None):
    """PIL.Image -> (Tensor[C, H, W] RGB in [-1, 1], h0, w0) — SUPIR/util.py:60-84 (the keyword `upsacle` is the reference's)."""
    from PIL import Image
    (w, h), (w0, h0) = _target_size(img.size, upsacle, min_size, fix_resize)
    pixels = np.array(img.resize((w, h), Image.BICUBIC)).round().clip(0, 255).astype(np.uint8)
    x = torch.tensor(pixels / 255 * 2 - 1, dtype=torch.float32).permute(2, 0, 1)
    return x, h0, w0


def tensor_to_uint8(x, h0, w0):
    """fp32 CUDA tensor [C, H, W] in [-1, 1] -> uint8 CUDA tensor [h0, w0, C] (the arithmetic of Tensor2PIL)."""
    _need_cuda(x)
    x = x.contiguous().float()
    C, H, W = x.shape
    out = torch.empty((h0, w0, C), dtype=torch.uint8, device=x.device)
    call("supir_image_to_uint8_bicubic", _ptr(x), C, H, W, _ptr(out), int(h0), int(w0), _stream())
    return out


def Tensor2PIL(x, h0, w0):
    """Tensor[C, H, W], RGB, [-1, 1] -> PIL.Image (SUPIR/util.py:87-94)."""
    from PIL import Image
    return Image.fromarray(tensor_to_uint8(x, h0, w0).cpu().numpy())
