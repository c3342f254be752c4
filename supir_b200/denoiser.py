"""Sigma schedule + Eps-preconditioned discrete denoiser (reference: sgm/modules/diffusionmodules/denoiser.py:31-73,
denoiser_scaling.py:16-22, discretizer.py:42-69, util.py:19-32).

The sigma tables are tiny host-side constants (float64 numpy -> float32), computed exactly as the reference does so the
nearest-sigma quantisation is bit-identical. `DiscreteDenoiserWithControl.__call__` keeps the reference signature; its
elementwise arithmetic runs in the fused CUDA kernels of supir_b200/csrc/elementwise.cu.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import instantiate_from_config


class EpsWeighting:
    def __call__(self, sigma):
        return sigma ** -2.0


class EpsScaling:
    """c_skip = 1, c_out = -sigma, c_in = 1/sqrt(sigma^2+1), c_noise = sigma."""

    def __call__(self, sigma):
        return torch.ones_like(sigma), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class LegacyDDPMDiscretization:
    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64, device="cpu") ** 2
        self.alphas_cumprod = np.cumprod(1.0 - betas.numpy(), axis=0)

    def get_sigmas(self, n, device="cpu"):
        if n < self.num_timesteps:
            steps = np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
            ac = self.alphas_cumprod[steps]
        elif n == self.num_timesteps:
            ac = self.alphas_cumprod
        else:
            raise ValueError
        sigmas = torch.tensor((1 - ac) / ac, dtype=torch.float32, device=device) ** 0.5
        return torch.flip(sigmas, (0,))

    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sigmas = self.get_sigmas(n, device=device)
        if do_append_zero:
            sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
        return sigmas if not flip else torch.flip(sigmas, (0,))


class DiscreteDenoiserWithControl(nn.Module):
    def __init__(self, weighting_config, scaling_config, num_idx, discretization_config, do_append_zero=False,
                 quantize_c_noise=True, flip=True):
        super().__init__()
        self.weighting = instantiate_from_config(weighting_config)
        self.scaling = instantiate_from_config(scaling_config)
        if not isinstance(self.scaling, EpsScaling):
            raise NotImplementedError("only EpsScaling is on SUPIR's sampling path")
        sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise
        self._host_table = sigmas.detach().cpu().numpy().copy()
        # `sigmas` is a persistent buffer (denoiser.py:43): a checkpoint's `denoiser.sigmas` replaces it, and the host copy follows
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._refresh_host_table())

    def _refresh_host_table(self):
        self._host_table = self.sigmas.detach().float().cpu().numpy().copy()

    def w(self, sigma):
        return self.weighting(sigma)

    # ---- host-side (exact) quantisation used by the fused sampler path ----
    def quantize_host(self, sigma: float):
        """(sigma_q, index) of the table entry nearest to sigma, with argmin's first-minimum tie rule (denoiser.py:49-51)."""
        idx = int(np.argmin(np.abs(np.float32(sigma) - self._host_table)))
        return float(self._host_table[idx]), idx

    @staticmethod
    def c_in_host(sigma_q: float) -> float:
        """EpsScaling's c_in = 1 / (sigma^2 + 1) ** 0.5 evaluated the way torch does on a float32 tensor (pow(x, 0.5) is sqrt):
        float32 square, add, correctly rounded sqrt, divide (denoiser_scaling.py:16-22). ONE definition, so the fused and the
        unfused sampler paths feed bit-identical inputs to the network."""
        s = np.float32(sigma_q)
        return float(np.float32(1.0) / np.sqrt(s * s + np.float32(1.0)))

    # ---- reference-compatible tensor API ----
    def sigma_to_idx(self, sigma):
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise

    @torch.no_grad()
    def __call__(self, network, input, sigma, cond, control_scale):
        """network(input * c_in, idx, cond, control_scale) * c_out + input  (denoiser.py:66-73).
        All samples of a call share one sigma in SUPIR's samplers; per-sample sigmas fall back to a loop over groups."""
        sig = sigma.detach().float().cpu().numpy().reshape(-1)
        if not np.all(sig == sig[0]):
            outs = [self(network, input[i:i + 1], sigma[i:i + 1], {k: (v[i:i + 1] if torch.is_tensor(v) else v) for k, v in cond.items()},
                         control_scale) for i in range(input.shape[0])]
            return torch.cat(outs, 0)
        sq, idx = self.quantize_host(float(sig[0]))
        c_in = self.c_in_host(sq)
        x = input.contiguous().float()
        scaled = torch.empty_like(x)
        ops.axpby_f32(x, float(c_in), None, 0.0, scaled)
        t = torch.full((x.shape[0],), idx, dtype=torch.long, device=x.device)
        net = network(scaled, t, cond, control_scale)
        out = torch.empty_like(x)
        ops.axpby_f32(net.contiguous(), -sq, x, 1.0, out)
        return out
