"""Image pre/post-processing around the sampler (reference: SUPIR/util.py:60-94).

`PIL2Tensor` is host-side input preparation (a PIL resize) and is kept as the reference has it. `Tensor2PIL` — bicubic
resize back to the requested size, scale to 0..255, clip, uint8, HWC — sits right after the final VAE decode and runs as one
CUDA kernel (supir_image_to_uint8_bicubic); only the uint8 image crosses PCIe."""
import numpy as np
import torch

from ._native import call
from .ops import _need_cuda, _ptr, _stream


def PIL2Tensor(img, upsacle=1, min_size=1024, fix_resize=None):
    """PIL.Image -> Tensor[C, H, W], RGB, [-1, 1] (SUPIR/util.py:60-84; argument names as in the reference)."""
    from PIL import Image
    w, h = img.size
    w *= upsacle
    h *= upsacle
    w0, h0 = round(w), round(h)
    if min(w, h) < min_size:
        _upsacle = min_size / min(w, h)
        w *= _upsacle
        h *= _upsacle
    if fix_resize is not None:
        _upsacle = fix_resize / min(w, h)
        w *= _upsacle
        h *= _upsacle
        w0, h0 = round(w), round(h)
    w = int(np.round(w / 64.0)) * 64
    h = int(np.round(h / 64.0)) * 64
    x = img.resize((w, h), Image.BICUBIC)
    x = np.array(x).round().clip(0, 255).astype(np.uint8)
    x = x / 255 * 2 - 1
    x = torch.tensor(x, dtype=torch.float32).permute(2, 0, 1)
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
