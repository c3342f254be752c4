"""Oracle: colour fix of the decoded image (SUPIR/utils/colorfix.py). Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F


def wavelet_blur(image, radius):
    """colorfix.py:73-92: depthwise 3x3 [1 2 1]x[1 2 1]/16 at dilation `radius`, replicate padding."""
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], dtype=image.dtype)
    c = image.shape[1]
    k = k[None, None].repeat(c, 1, 1, 1)
    image = F.pad(image, (radius, radius, radius, radius), mode="replicate")
    return F.conv2d(image, k, groups=c, dilation=radius)


def wavelet_decomposition(image, levels=5):
    """colorfix.py:94-106."""
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high = high + (image - low)
        image = low
    return high, low


def wavelet_reconstruction(content, style):
    """colorfix.py:108-120."""
    return wavelet_decomposition(content)[0] + wavelet_decomposition(style)[1]


def adaptive_instance_normalization(content, style, eps=1e-5):
    """colorfix.py:45-71 (unbiased variance + eps)."""
    def ms(f):
        b, c = f.shape[:2]
        var = f.reshape(b, c, -1).var(dim=2) + eps
        return f.reshape(b, c, -1).mean(dim=2).reshape(b, c, 1, 1), var.sqrt().reshape(b, c, 1, 1)
    sm, ss = ms(style)
    cm, cs = ms(content)
    return (content - cm) / cs * ss + sm
