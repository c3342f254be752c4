"""Classifier-free guidance: batch-pair build and reduction (reference: sgm/modules/diffusionmodules/guiders.py:7-88,
sampling_utils.py:7-9). The pair reduction is a CUDA kernel (supir_cfg_combine); the batch-pair build only moves handles."""
import torch

from . import ops

_PAIR_KEYS = ["vector", "crossattn", "concat", "control", "control_vector", "mask_x"]


class NoDynamicThresholding:
    def __call__(self, uncond, cond, scale):
        x = torch.cat([uncond, cond], 0).contiguous().float()
        out = torch.empty_like(uncond, dtype=torch.float32)
        ops.cfg_combine(x, scale.reshape(-1).float().contiguous(), out)
        return out


class _CFGBase:
    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in _PAIR_KEYS:
                c_out[k] = torch.cat((uc[k], c[k]), 0)       # unconditional first (guiders.py:65-74)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out

    def __call__(self, x, sigma):
        scale = self.scale_schedule(sigma)
        if not torch.is_tensor(scale):
            scale = torch.full((x.shape[0] // 2,), float(scale), device=x.device)
        out = torch.empty((x.shape[0] // 2,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        ops.cfg_combine(x.contiguous().float(), scale.reshape(-1).float().contiguous(), out)
        return out


class VanillaCFG(_CFGBase):
    def __init__(self, scale, dyn_thresh_config=None):
        self.scale = scale
        self.scale_min = scale

    def scale_schedule(self, sigma):
        return self.scale

    def scale_host(self, sigma: float):
        return float(self.scale)


class LinearCFG(_CFGBase):
    """scale(sigma) = (scale - scale_min) * sigma / 14.6146 + scale_min (guiders.py:44-63)."""

    def __init__(self, scale, scale_min=None, dyn_thresh_config=None):
        self.scale = scale
        self.scale_min = scale if scale_min is None else scale_min

    def scale_schedule(self, sigma):
        return (self.scale - self.scale_min) * sigma / 14.6146 + self.scale_min

    def scale_host(self, sigma: float):
        """Same expression evaluated like the reference does on a float32 sigma tensor."""
        import numpy as np
        return float(np.float32(self.scale - self.scale_min) * np.float32(sigma) / np.float32(14.6146) + np.float32(self.scale_min))


class IdentityGuider:
    def __call__(self, x, sigma):
        return x

    def prepare_inputs(self, x, s, c, uc):
        return x, s, dict(c)
