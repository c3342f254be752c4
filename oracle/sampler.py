"""Oracle: sigma schedule, Eps denoiser, LinearCFG, RestoreEDMSampler and TiledRestoreEDMSampler (fp32 torch / numpy).

Test infrastructure only (see oracle/__init__.py). `network(x, t_idx, cond, control_scale)` is any callable with the
ControlWrapper signature (sgm/modules/diffusionmodules/wrappers.py:84-102).
"""
import numpy as np
import torch

SIGMA_MAX = 14.6146  # hard-coded in the reference (sampling.py:541, guiders.py:48)


def legacy_ddpm_alphas_cumprod(linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
    """make_beta_schedule('linear') + cumprod (sgm/modules/diffusionmodules/util.py:19-32; discretizer.py:42-55)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2
    return np.cumprod(1.0 - betas.numpy(), axis=0)


def legacy_ddpm_sigmas(n, do_append_zero=True, flip=False):
    """LegacyDDPMDiscretization.get_sigmas + Discretization.__call__ (discretizer.py:17-21, 57-69)."""
    ac = legacy_ddpm_alphas_cumprod()
    if n < 1000:
        timesteps = np.linspace(1000 - 1, 0, n, endpoint=False).astype(int)[::-1]  # discretizer.py:11-14
        ac = ac[timesteps]
    elif n != 1000:
        raise ValueError
    sigmas = torch.tensor((1 - ac) / ac, dtype=torch.float32) ** 0.5
    sigmas = torch.flip(sigmas, (0,))
    if do_append_zero:
        sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
    return sigmas if not flip else torch.flip(sigmas, (0,))


def denoiser_sigma_table():
    """DiscreteDenoiser.__init__ (denoiser.py:31-47): 1000 sigmas, ascending (flip=True), no appended zero."""
    return legacy_ddpm_sigmas(1000, do_append_zero=False, flip=True)


def sigma_to_idx(table, sigma):
    """DiscreteDenoiser.sigma_to_idx (denoiser.py:49-51)."""
    dists = sigma - table[:, None]
    return dists.abs().argmin(dim=0).view(sigma.shape)


def denoise_with_control(network, table, x, sigma, cond, control_scale):
    """DiscreteDenoiserWithControl.__call__ (denoiser.py:66-73) with EpsScaling (denoiser_scaling.py:16-22)."""
    sigma = table[sigma_to_idx(table, sigma)]
    sigma_shape = sigma.shape
    s = sigma[(...,) + (None,) * (x.ndim - sigma.ndim)]
    c_skip, c_out, c_in, c_noise = torch.ones_like(s), -s, 1 / (s ** 2 + 1.0) ** 0.5, s.clone()
    c_noise = sigma_to_idx(table, c_noise.reshape(sigma_shape))
    return network(x * c_in, c_noise, cond, control_scale) * c_out + x * c_skip


def cfg_prepare_inputs(x, s, c, uc):
    """LinearCFG.prepare_inputs (guiders.py:65-74): unconditional first."""
    c_out = {}
    for k in c:
        if k in ["vector", "crossattn", "concat", "control", "control_vector", "mask_x"]:
            c_out[k] = torch.cat((uc[k], c[k]), 0)
        else:
            c_out[k] = c[k]
    return torch.cat([x] * 2), torch.cat([s] * 2), c_out


def linear_cfg_scale(scale, scale_min, sigma):
    """LinearCFG scale_schedule (guiders.py:48)."""
    return (scale - scale_min) * sigma / SIGMA_MAX + scale_min


def cfg_combine(x, sigma, scale, scale_min):
    """LinearCFG.__call__ + NoDynamicThresholding (guiders.py:59-63; sampling_utils.py:7-9)."""
    x_u, x_c = x.chunk(2)
    s = linear_cfg_scale(scale, scale_min, sigma)
    return x_u + s.view(-1, 1, 1, 1) * (x_c - x_u)


class RestoreEDMSampler:
    """sampling.py:528-597. `randn_like` is injectable so tests can feed the same noise to both implementations."""

    def __init__(self, num_steps, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, restore_cfg=4.0,
                 restore_cfg_s_tmin=0.05, scale=7.5, scale_min=4.0, randn_like=torch.randn_like):
        self.num_steps, self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = num_steps, s_churn, s_tmin, s_tmax, s_noise
        self.restore_cfg, self.restore_cfg_s_tmin = restore_cfg, restore_cfg_s_tmin
        self.scale, self.scale_min = scale, scale_min
        self.randn_like = randn_like
        self.table = denoiser_sigma_table()

    def denoise(self, network, x, sigma, cond, uc, control_scale):
        xi, si, ci = cfg_prepare_inputs(x, sigma, cond, uc)
        den = denoise_with_control(network, self.table, xi, si, ci, control_scale)
        return cfg_combine(den, sigma, self.scale, self.scale_min)

    def sampler_step(self, network, sigma, next_sigma, x, cond, uc, gamma, x_center, eps_noise=None, control_scale=1.0,
                     use_linear_control_scale=False, control_scale_start=0.0):
        """sampling.py:548-570."""
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = (eps_noise if eps_noise is not None else self.randn_like(x)) * self.s_noise
            x = x + eps * ((sigma_hat ** 2 - sigma ** 2)[:, None, None, None]) ** 0.5
        if use_linear_control_scale:
            control_scale = (sigma[0].item() / SIGMA_MAX) * (control_scale_start - control_scale) + control_scale
        denoised = self.denoise(network, x, sigma_hat, cond, uc, control_scale)
        if (next_sigma[0] > self.restore_cfg_s_tmin) and (self.restore_cfg > 0):
            d_center = denoised - x_center
            denoised = denoised - d_center * ((sigma.view(-1, 1, 1, 1) / SIGMA_MAX) ** self.restore_cfg)
        d = (x - denoised) / sigma_hat[:, None, None, None]
        dt = (next_sigma - sigma_hat)[:, None, None, None]
        return x + dt * d  # euler_step (sampling.py:121-122)

    def prepare(self, x):
        """prepare_sampling_loop (sampling.py:45-56)."""
        sigmas = legacy_ddpm_sigmas(self.num_steps)
        x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
        return x, x.new_ones([x.shape[0]]), sigmas

    def gamma(self, sigmas, i):
        n = len(sigmas) - 1
        return min(self.s_churn / n, 2 ** 0.5 - 1) if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0

    def __call__(self, network, x, cond, uc, x_center, control_scale=1.0, use_linear_control_scale=False,
                 control_scale_start=0.0):
        x, s_in, sigmas = self.prepare(x)
        for i in range(len(sigmas) - 1):
            x = self.sampler_step(network, s_in * sigmas[i], s_in * sigmas[i + 1], x, cond, uc, self.gamma(sigmas, i),
                                  x_center, control_scale=control_scale,
                                  use_linear_control_scale=use_linear_control_scale,
                                  control_scale_start=control_scale_start)
        return x


def gaussian_weights(tile_width, tile_height):
    """sampling.py:733-750 — float64; note the x midpoint is (W-1)/2 but the y midpoint is H/2 (reference quirk)."""
    var = 0.01
    mid = (tile_width - 1) / 2
    x_probs = [np.exp(-(x - mid) * (x - mid) / (tile_width * tile_width) / (2 * var)) / np.sqrt(2 * np.pi * var)
               for x in range(tile_width)]
    mid = tile_height / 2
    y_probs = [np.exp(-(y - mid) * (y - mid) / (tile_height * tile_height) / (2 * var)) / np.sqrt(2 * np.pi * var)
               for y in range(tile_height)]
    return np.outer(y_probs, x_probs)  # float64 [H, W]


def sliding_windows(h, w, tile_size, tile_stride):
    """sampling.py:753-766."""
    hi_list = list(range(0, h - tile_size + 1, tile_stride))
    if (h - tile_size) % tile_stride != 0:
        hi_list.append(h - tile_size)
    wi_list = list(range(0, w - tile_size + 1, tile_stride))
    if (w - tile_size) % tile_stride != 0:
        wi_list.append(w - tile_size)
    return [(hi, hi + tile_size, wi, wi + tile_size) for hi in hi_list for wi in wi_list]


class TiledRestoreEDMSampler(RestoreEDMSampler):
    """sampling.py:600-660."""

    def __init__(self, tile_size=128, tile_stride=64, **kw):
        super().__init__(**kw)
        self.tile_size, self.tile_stride = tile_size, tile_stride
        self.tile_weights = torch.tensor(gaussian_weights(tile_size, tile_size)).repeat(1, 4, 1, 1)  # float64

    def __call__(self, network, x, cond, uc, x_center, control_scale=1.0, use_linear_control_scale=False,
                 control_scale_start=0.0):
        b, _, h, w = x.shape
        windows = sliding_windows(h, w, self.tile_size, self.tile_stride)
        tile_weights = self.tile_weights.repeat(b, 1, 1, 1)
        use_local_prompt = isinstance(cond, list)          # one conditioning dict per window (sampling.py:609-617)
        if use_local_prompt:
            assert len(cond) == len(windows)
            cond = [dict(c) for c in cond]
            lq = cond[0]["control"]
        else:
            cond = dict(cond)
            lq = cond["control"]
        uc = dict(uc)
        x, s_in, sigmas = self.prepare(x)
        for i in range(len(sigmas) - 1):
            gamma = self.gamma(sigmas, i)
            x_next = torch.zeros_like(x)
            count = torch.zeros_like(x)
            eps_noise = self.randn_like(x)
            for j, (hi, he, wi, we) in enumerate(windows):
                _cond = cond[j] if use_local_prompt else cond
                _cond["control"] = lq[:, :, hi:he, wi:we]
                uc["control"] = lq[:, :, hi:he, wi:we]
                _x = self.sampler_step(network, s_in * sigmas[i], s_in * sigmas[i + 1], x[:, :, hi:he, wi:we], _cond, uc,
                                       gamma, x_center[:, :, hi:he, wi:we], eps_noise=eps_noise[:, :, hi:he, wi:we],
                                       control_scale=control_scale, use_linear_control_scale=use_linear_control_scale,
                                       control_scale_start=control_scale_start)
                x_next[:, :, hi:he, wi:we] += _x * tile_weights
                count[:, :, hi:he, wi:we] += tile_weights
            x_next /= count
            x = x_next
        return x


# --------------------------------------------------------------------------------------------------------------------
# DPM++ 2M SDE restore samplers (Lightning configs) — sampling.py:271-360 (DPMPP2MSampler), :422-515, :663-730
# --------------------------------------------------------------------------------------------------------------------
def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0):
    """k-diffusion 0.1.1 `sampling.get_sigmas_karras` (requirements.txt:41; not vendored in the reference): Karras et al.
    (2022) schedule with an appended zero. Restated from the published formula — parity UNPINNED for this function."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


class RestoreDPMPP2MSampler:
    """sampling.py:422-515. `noise_sampler(sigma, sigma_next)` replaces k-diffusion's BrownianTreeNoiseSampler (torchsde is not
    available): any callable returning unit-variance noise of x's shape."""

    def __init__(self, num_steps, s_noise=1.0, eta=1.0, scale=7.5, scale_min=4.0, noise_sampler=None):
        self.num_steps, self.s_noise, self.eta = num_steps, s_noise, eta
        self.scale, self.scale_min = scale, scale_min
        self.noise_sampler = noise_sampler
        self.table = denoiser_sigma_table()

    def denoise(self, network, x, sigma, cond, uc, control_scale):
        xi, si, ci = cfg_prepare_inputs(x, sigma, cond, uc)
        den = denoise_with_control(network, self.table, xi, si, ci, control_scale)
        return cfg_combine(den, sigma, self.scale, self.scale_min)

    def sampler_step(self, network, old_denoised, previous_sigma, sigma, next_sigma, x, cond, uc, eps_noise, control_scale):
        denoised = self.denoise(network, x, sigma, cond, uc, control_scale)
        t, t_next = sigma.log().neg(), next_sigma.log().neg()
        h = t_next - t
        eta_h = self.eta * h
        d = lambda v: v[:, None, None, None]  # noqa: E731
        mult1 = t_next.neg().exp() / t.neg().exp() * (-eta_h).exp()
        mult2 = (-h - eta_h).expm1()
        x_standard = d(mult1) * x - d(mult2) * denoised
        if old_denoised is None or torch.sum(next_sigma) < 1e-14:
            return x_standard, denoised
        r = (t - previous_sigma.log().neg()) / h
        denoised_d = d(1 + 1 / (2 * r)) * denoised - d(1 / (2 * r)) * old_denoised
        x_advanced = d(mult1) * x - d(mult2) * denoised_d
        x = torch.where(d(next_sigma) > 0.0, x_advanced, x_standard)
        if self.eta:
            x = x + eps_noise * d(next_sigma) * d((-2 * eta_h).expm1().neg().sqrt()) * self.s_noise
        return x, denoised

    def schedule(self, x):
        sig = legacy_ddpm_sigmas(self.num_steps)
        x = x * torch.sqrt(1.0 + sig[0] ** 2.0)
        return x, x.new_ones([x.shape[0]]), get_sigmas_karras(self.num_steps, sig[-2], sig[0]), len(sig)

    def __call__(self, network, x, cond, uc, control_scale=1.0):
        x, s_in, sigmas, num_sigmas = self.schedule(x)
        old = None
        for i in range(num_sigmas - 1):
            eps = None
            if i > 0 and torch.sum(s_in * sigmas[i + 1]) > 1e-14:
                eps = self.noise_sampler(s_in * sigmas[i], s_in * sigmas[i + 1])
            x, old = self.sampler_step(network, old, None if i == 0 else s_in * sigmas[i - 1], s_in * sigmas[i],
                                       s_in * sigmas[i + 1], x, cond, uc, eps, control_scale)
        return x


class TiledRestoreDPMPP2MSampler(RestoreDPMPP2MSampler):
    """sampling.py:663-730."""

    def __init__(self, tile_size=128, tile_stride=64, **kw):
        super().__init__(**kw)
        self.tile_size, self.tile_stride = tile_size, tile_stride
        self.tile_weights = torch.tensor(gaussian_weights(tile_size, tile_size)).repeat(1, 4, 1, 1)

    def __call__(self, network, x, cond, uc, control_scale=1.0):
        b, _, h, w = x.shape
        windows = sliding_windows(h, w, self.tile_size, self.tile_stride)
        tw = self.tile_weights.repeat(b, 1, 1, 1)
        lq = cond["control"]
        x, s_in, sigmas, num_sigmas = self.schedule(x)
        cond, uc = dict(cond), dict(uc)
        old = None
        for i in range(num_sigmas - 1):
            if i > 0 and torch.sum(s_in * sigmas[i + 1]) > 1e-14:
                eps_noise = self.noise_sampler(s_in * sigmas[i], s_in * sigmas[i + 1])
            else:
                eps_noise = torch.zeros_like(x)
            x_next, old_next, count = torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)
            for (hi, he, wi, we) in windows:
                cond["control"] = lq[:, :, hi:he, wi:we]
                uc["control"] = lq[:, :, hi:he, wi:we]
                _x, _old = self.sampler_step(network, None if old is None else old[:, :, hi:he, wi:we],
                                             None if i == 0 else s_in * sigmas[i - 1], s_in * sigmas[i], s_in * sigmas[i + 1],
                                             x[:, :, hi:he, wi:we], cond, uc, eps_noise[:, :, hi:he, wi:we], control_scale)
                x_next[:, :, hi:he, wi:we] += _x * tw
                old_next[:, :, hi:he, wi:we] += _old * tw
                count[:, :, hi:he, wi:we] += tw
            old_next /= count
            x_next /= count
            x, old = x_next, old_next
        return x
