"""EDM Euler samplers of SUPIR (reference: sgm/modules/diffusionmodules/sampling.py:25-75, 528-660, 733-766).

`RestoreEDMSampler` / `TiledRestoreEDMSampler` keep the reference constructor and call signatures. What differs is HOW a
step runs:
  * every sigma-dependent scalar (gamma, sigma_hat, c_in, c_out, CFG scale, restore factor, dt, linear control scale) is
    computed on the host in float32 from the schedule, so a step has NO device->host sync (the reference has >= 2);
  * the per-step latent arithmetic is two fused kernels (supir_edm_pre / supir_edm_post) around one CUDA-graph replay of
    the control + UNet pair, when the denoiser is this package's (`FusedDenoiser`); a generic callable still works through
    the unfused kernels;
  * the tiled sampler stacks several windows into one network batch, and, under torch.distributed, shards the windows
    over ranks with ONE all-gather of the window outputs per step; the Gaussian blend then runs on every rank in the
    reference's window order (bit-identical x_{t+1} on all ranks; sampling.py:629-659).
RNG stays in PyTorch (torch.randn_like in the reference's order); kernels take the noise as an input.
"""
import numpy as np
import torch

from . import ops
from .brownian import BrownianTreeNoiseSampler
from .config import instantiate_from_config

import itertools

_RUN_IDS = itertools.count(1)      # names a sampler run (its conditioning is constant) for ControlWrapper's text K|V reuse
SIGMA_MAX = 14.6146
f32 = np.float32
DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


def gaussian_weights(tile_width, tile_height, nbatches, device="cuda"):
    """sampling.py:733-750 (float64; x midpoint (W-1)/2, y midpoint H/2 — the reference's asymmetry is kept)."""
    from numpy import exp, pi, sqrt
    var = 0.01
    midpoint = (tile_width - 1) / 2
    x_probs = [exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) / sqrt(2 * pi * var)
               for x in range(tile_width)]
    midpoint = tile_height / 2
    y_probs = [exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) / sqrt(2 * pi * var)
               for y in range(tile_height)]
    weights = np.outer(y_probs, x_probs)
    return torch.tile(torch.tensor(weights, device=device), (nbatches, 4, 1, 1))


def _sliding_windows(h: int, w: int, tile_size: int, tile_stride: int):
    """sampling.py:753-766: (hi, hi_end, wi, wi_end) per window, row-major, last window clamped to the border."""
    def starts(n):
        s = list(range(0, n - tile_size + 1, tile_stride))
        if (n - tile_size) % tile_stride != 0:
            s.append(n - tile_size)
        return s
    return [(hi, hi + tile_size, wi, wi + tile_size) for hi in starts(h) for wi in starts(w)]


def shard_windows(num_windows, world_size, rank):
    """Contiguous block partition: rank r owns slots [r*per, (r+1)*per) of a table padded to per*world_size slots.
    Contiguity keeps the slot order equal to the reference's window order, which the ordered blend relies on."""
    per = (num_windows + world_size - 1) // world_size
    lo = min(rank * per, num_windows)
    hi = min(lo + per, num_windows)
    return per, lo, hi


def shard_units(num_units, world_size, rank):
    """Balanced contiguous partition of the (CFG branch, window) units of a tiled step: sizes differ by at most one
    (98 units over 8 ranks -> 13, 13, 12 x 6 instead of the 14 / 0 a per-window split leaves). Returns (slot size of the
    padded exchange buffer, first unit, one past the last unit)."""
    base, rem = divmod(num_units, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return base + (1 if rem else 0), lo, hi


def balanced_groups(n, max_group):
    """Split n windows into ceil(n / max_group) contiguous groups whose sizes differ by at most one (at most two
    distinct network batch sizes -> at most two captured CUDA graphs)."""
    if n <= 0:
        return []
    k = (n + max_group - 1) // max_group
    base, rem = divmod(n, k)
    out, lo = [], 0
    for i in range(k):
        hi = lo + base + (1 if i < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exchange_window_outputs(tiles_out, rank, per, group=None):
    """The one collective of a tiled step on the generic path: in-place all-gather of the per-rank blocks of `tiles_out`
    ([world*per, ...] fp32, rank r owns slots [r*per, (r+1)*per)). NCCL on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    mine = tiles_out[rank * per:(rank + 1) * per].reshape(-1).clone()
    dist.all_gather_into_tensor(tiles_out.view(-1), mine, group=group)
    return tiles_out


def exchange_unit_outputs(padded, rank, per, group=None):
    """The one collective of a tiled step on the fused path: in-place all-gather of the per-rank slots of `padded`
    ([world*per, ...] fp32, rank r's units in slots [r*per, r*per + n_r))."""
    import torch.distributed as dist
    mine = padded[rank * per:(rank + 1) * per].reshape(-1).clone()
    dist.all_gather_into_tensor(padded.view(-1), mine, group=group)
    return padded


class FusedDenoiser:
    """Callable handed to the samplers by supir_b200.model.SUPIRModel: behaves like the reference's lambda
    (SUPIR_model.py:123-125) and additionally exposes the parts so the sampler can fuse the step arithmetic."""

    def __init__(self, denoiser, network):
        self.denoiser, self.network = denoiser, network
        self._tokens = bool(getattr(network, "supports_context_token", False))

    def __call__(self, input, sigma, c, control_scale):
        return self.denoiser(self.network, input, sigma, c, control_scale)

    def run_network(self, x, t, c, control_scale, context_token=None):
        """network(x, t, c, control_scale); the token (a name for the content of c['crossattn'], constant over a sampler run)
        is forwarded only to networks that understand it (supir_b200.wrappers.ControlWrapper)."""
        if self._tokens and context_token is not None:
            return self.network(x, t, c, control_scale, context_token=context_token)
        return self.network(x, t, c, control_scale)


class BaseDiffusionSampler:
    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False, device="cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(guider_config if guider_config is not None else DEFAULT_GUIDER)
        self.verbose = verbose
        self.device = device

    def host_sigmas(self, num_steps=None):
        s = self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu")
        return s.numpy().astype(np.float32)


class RestoreEDMSampler(BaseDiffusionSampler):
    shard = False            # opt-in: run the two CFG branches of an untiled step on two torch.distributed ranks (_EDMRun)
    process_group = None

    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, restore_cfg=4.0, restore_cfg_s_tmin=0.05,
                 *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise
        self.restore_cfg, self.restore_cfg_s_tmin = restore_cfg, restore_cfg_s_tmin
        self.sigma_max = SIGMA_MAX

    # ---- host-side step constants, evaluated in float32 exactly where the reference uses float32 tensors ----
    def step_constants(self, sigmas, i, control_scale, use_linear_control_scale, control_scale_start):
        n = len(sigmas) - 1
        sigma, next_sigma = f32(sigmas[i]), f32(sigmas[i + 1])
        gamma = min(self.s_churn / n, 2 ** 0.5 - 1) if self.s_tmin <= sigma <= self.s_tmax else 0.0
        sigma_hat = f32(sigma * f32(gamma + 1.0))
        noise_mul = 0.0
        if gamma > 0:
            noise_mul = float(f32(self.s_noise) * f32(np.sqrt(f32(sigma_hat * sigma_hat - sigma * sigma))))
        cs = control_scale
        if use_linear_control_scale:
            cs = (float(sigma) / self.sigma_max) * (control_scale_start - control_scale) + control_scale
        restore_mul = 0.0
        use_restore = (next_sigma > self.restore_cfg_s_tmin) and (self.restore_cfg > 0)
        if use_restore:
            restore_mul = float(f32(sigma / f32(self.sigma_max)) ** f32(self.restore_cfg))
        return dict(sigma=float(sigma), next_sigma=float(next_sigma), gamma=gamma, sigma_hat=float(sigma_hat),
                    noise_mul=noise_mul, control_scale=cs, use_restore=bool(use_restore), restore_mul=restore_mul,
                    dt=float(f32(next_sigma - sigma_hat)))

    # ---- one step on a batch of latents (x: fp32 [N,4,h,w]); eps may be None when gamma == 0 ----
    def _step(self, denoiser, x, eps, cond, uc, x_center, k):
        N = x.shape[0]
        sig_hat_t = torch.full((N,), k["sigma_hat"], dtype=torch.float32, device=x.device)
        if isinstance(denoiser, FusedDenoiser) and hasattr(self.guider, "scale_host"):
            den = denoiser.denoiser
            sq, idx = den.quantize_host(k["sigma_hat"])
            c_in = den.c_in_host(sq)
            x_hat = torch.empty_like(x)
            net_in = torch.empty((2 * N,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
            ops.edm_pre(x, eps if k["gamma"] > 0 else None, k["noise_mul"], c_in, x_hat, net_in)
            cpair = {key: torch.cat((uc[key], cond[key]), 0) for key in ("vector", "crossattn", "control")}
            t = torch.full((2 * N,), idx, dtype=torch.long, device=x.device)
            net_out = denoiser.run_network(net_in, t, cpair, k["control_scale"], k.get("context_token"))
            x_next = torch.empty_like(x)
            ops.edm_post(x_hat, net_out, x_center if k["use_restore"] else None, -sq, self.guider.scale_host(k["sigma_hat"]),
                         k["restore_mul"], k["sigma_hat"], k["dt"], x_next)
            return x_next
        # generic callable (e.g. the reference's own denoiser lambda): unfused kernels, same arithmetic
        x_hat = x
        if k["gamma"] > 0:
            x_hat = torch.empty_like(x)
            ops.axpby_f32(x, 1.0, eps, k["noise_mul"], x_hat)
        xi, si, ci = self.guider.prepare_inputs(x_hat, sig_hat_t, cond, uc)
        denoised = self.guider(denoiser(xi, si, ci, k["control_scale"]), sig_hat_t).contiguous()
        if k["use_restore"]:
            d2 = torch.empty_like(denoised)
            ops.axpby_f32(denoised, 1.0 - k["restore_mul"], x_center.contiguous(), k["restore_mul"], d2)
            denoised = d2
        r = k["dt"] / k["sigma_hat"]
        x_next = torch.empty_like(x)
        ops.axpby_f32(x_hat.contiguous(), 1.0 + r, denoised, -r, x_next)
        return x_next

    def prepare_sampling_loop(self, x, num_steps=None):
        sigmas = self.host_sigmas(num_steps)
        x0 = torch.empty_like(x, dtype=torch.float32)
        ops.axpby_f32(x.contiguous().float(), float(np.sqrt(f32(1.0) + sigmas[0] * sigmas[0])), None, 0.0, x0)
        return x0, sigmas

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, x_center=None, control_scale=1.0,
                 use_linear_control_scale=False, control_scale_start=0.0):
        run = self.begin(denoiser, x, cond, uc, num_steps, x_center, control_scale, use_linear_control_scale, control_scale_start)
        for i in range(run.num_steps):
            run.step(i)
        return run.x

    @torch.no_grad()
    def begin(self, denoiser, x, cond, uc=None, num_steps=None, x_center=None, control_scale=1.0,
              use_linear_control_scale=False, control_scale_start=0.0):
        """Set up a run and return an object whose .step(i) advances one EDM step (bench.py times those)."""
        return _EDMRun(self, denoiser, x, cond, cond if uc is None else uc, num_steps, x_center, control_scale,
                       use_linear_control_scale, control_scale_start)


class _EDMRun:
    """One untiled sampling run. With `smp.shard` set on a 2-rank group (opt-in, identical inputs and RNG state on both
    ranks) the two CFG branches of every step run on different GPUs — rank 0 the unconditional rows, rank 1 the
    conditional ones (halves the latency of an untiled image) — with ONE all-gather of the raw
    network outputs per step; both ranks then finish the step, so x stays bit-identical on both and equal to the
    single-GPU result."""

    def __init__(self, smp, denoiser, x, cond, uc, num_steps, x_center, control_scale, lin_cs, cs_start):
        import torch.distributed as dist
        self.smp, self.denoiser, self.cond, self.uc = smp, denoiser, cond, uc
        self.uid = next(_RUN_IDS)
        self.group = getattr(smp, "process_group", None)
        fused = isinstance(denoiser, FusedDenoiser) and hasattr(smp.guider, "scale_host")
        self.world, self.rank = ((dist.get_world_size(self.group), dist.get_rank(self.group))
                                 if getattr(smp, "shard", False) and fused and dist.is_available() and dist.is_initialized() else (1, 0))
        if self.world > 1:
            self.per, self.lo, self.hi = shard_units(2, self.world, self.rank)       # units = the two CFG branches
            self.pair_out = torch.zeros((self.per * self.world,) + tuple(x.shape), dtype=torch.float32, device=x.device)
        self.control_scale, self.lin_cs, self.cs_start = control_scale, lin_cs, cs_start
        self.x, self.sigmas = smp.prepare_sampling_loop(x, num_steps)
        self.xc = None if x_center is None else x_center.contiguous().float()
        self.num_steps = len(self.sigmas) - 1

    @torch.no_grad()
    def step(self, i):
        k = self.smp.step_constants(self.sigmas, i, self.control_scale, self.lin_cs, self.cs_start)
        k["context_token"] = self.uid          # cond / uc are this run's, unchanged over its steps
        eps = torch.randn_like(self.x) if k["gamma"] > 0 else None
        if self.world > 1:
            self.x = self._step_branch_parallel(k, eps)
        else:
            self.x = self.smp._step(self.denoiser, self.x, eps, self.cond, self.uc, self.xc, k)
        return self.x

    def _step_branch_parallel(self, k, eps):
        smp, x = self.smp, self.x
        N = x.shape[0]
        sq, idx = self.denoiser.denoiser.quantize_host(k["sigma_hat"])
        c_in = self.denoiser.denoiser.c_in_host(sq)
        x_hat = torch.empty_like(x)
        net_in = torch.empty((2 * N,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        ops.edm_pre(x, eps if k["gamma"] > 0 else None, k["noise_mul"], c_in, x_hat, net_in)
        for u in range(self.lo, self.hi):            # unit 0: unconditional rows, unit 1: conditional rows
            c = self.uc if u == 0 else self.cond
            t = torch.full((N,), idx, dtype=torch.long, device=x.device)
            out = self.denoiser.run_network(net_in[u * N:(u + 1) * N], t, {key: c[key] for key in ("vector", "crossattn", "control")},
                                            k["control_scale"], (self.uid, u))
            self.pair_out[self.rank * self.per + (u - self.lo)] = out
        exchange_unit_outputs(self.pair_out, self.rank, self.per, self.group)
        if self.per * self.world == 2:
            both = self.pair_out
        else:                                         # more than two ranks: the first two carry one branch each
            both = torch.stack([self.pair_out[r * self.per] for r, (_, lo, hi) in
                                enumerate(shard_units(2, self.world, q) for q in range(self.world)) if hi > lo], 0)
        x_next = torch.empty_like(x)
        ops.edm_post(x_hat, both.reshape((2 * N,) + tuple(x.shape[1:])), self.xc if k["use_restore"] else None, -sq,
                     smp.guider.scale_host(k["sigma_hat"]), k["restore_mul"], k["sigma_hat"], k["dt"], x_next)
        return x_next


class TiledRestoreEDMSampler(RestoreEDMSampler):
    def __init__(self, tile_size=128, tile_stride=64, tile_batch=8, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.tile_size, self.tile_stride = tile_size, tile_stride
        self.tile_batch = max(1, int(tile_batch))        # windows stacked into one network call
        # Sharding the windows of ONE image over torch.distributed ranks is opt-in (a data-parallel caller with a different
        # image per rank must not meet collectives): set `.shard = True` (and optionally `.process_group`) on every rank,
        # with identical inputs and RNG state on all of them (SUPIRModel.enable_tile_sharding does this).
        self.shard = False
        self.process_group = None
        self._weights = None

    @property
    def tile_weights(self):
        if self._weights is None:
            self._weights = gaussian_weights(self.tile_size, self.tile_size, 1, device=self.device)
        return self._weights

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, x_center=None, control_scale=1.0,
                 use_linear_control_scale=False, control_scale_start=0.0):
        run = self.begin(denoiser, x, cond, uc, num_steps, x_center, control_scale, use_linear_control_scale,
                         control_scale_start)
        for i in range(run.num_steps):
            run.step(i)
        return run.x

    @torch.no_grad()
    def begin(self, denoiser, x, cond, uc=None, num_steps=None, x_center=None, control_scale=1.0,
              use_linear_control_scale=False, control_scale_start=0.0):
        """Set up a tiled run and return an object whose .step(i) advances one EDM step (bench.py times those)."""
        return _TiledRun(self, denoiser, x, cond, uc, num_steps, x_center, control_scale, use_linear_control_scale,
                         control_scale_start)


class _TiledRun:
    """One tiled sampling run (sampling.py:600-660). Two execution paths:

    fused (this package's FusedDenoiser + a guider with a host-side scale): the step arithmetic of ALL windows is two
      kernels (edm_pre / edm_post) around the network calls, and the work items handed to the network are the 2 x windows
      (CFG branch, window) UNITS in branch-major order (all unconditional rows, then all conditional rows — the layout
      edm_post reduces). Under sharding a rank runs a balanced contiguous range of units (they need not be pairs: both
      branches of a window see the same input), ONE all-gather moves the raw network outputs, and every rank finishes the
      step (edm_post + ordered Gaussian blend) on all windows -> bit-identical x on every rank and identical to the
      single-GPU run, because every kernel treats batch rows independently.
    generic (any callable denoiser, e.g. the reference's own lambda): windows are sharded as CFG pairs and the per-window
      step results are exchanged (the round-1 path)."""

    def __init__(self, smp, denoiser, x, cond, uc, num_steps, x_center, control_scale, use_linear_control_scale,
                 control_scale_start):
        import torch.distributed as dist
        self.smp, self.denoiser, self.cond = smp, denoiser, cond
        self.uid = next(_RUN_IDS)
        self.control_scale, self.lin_cs, self.cs_start = control_scale, use_linear_control_scale, control_scale_start
        self.use_local_prompt = isinstance(cond, list)
        b, ch, h, w = x.shape
        T = smp.tile_size
        windows = _sliding_windows(h, w, T, smp.tile_stride)
        nw = self.nw = len(windows)
        if self.use_local_prompt:
            assert len(cond) == nw, "Number of local prompts should be equal to number of tiles"
            lq = cond[0]["control"]
        else:
            lq = cond["control"]
        self.uc = (cond[0] if self.use_local_prompt else cond) if uc is None else uc
        lq = lq.contiguous().float()
        xc = x_center.contiguous().float()
        self.group = smp.process_group
        self.world, self.rank = ((dist.get_world_size(self.group), dist.get_rank(self.group))
                                 if smp.shard and dist.is_available() and dist.is_initialized() else (1, 0))
        self.fused = isinstance(denoiser, FusedDenoiser) and hasattr(smp.guider, "scale_host")
        self.weights64 = smp.tile_weights[0, 0].contiguous()
        self.x, self.sigmas = smp.prepare_sampling_loop(x, num_steps)
        self.num_steps = len(self.sigmas) - 1
        self.shape = (b, ch, T)
        dev = x.device
        if self.fused:
            self.table_dev = torch.tensor(windows, dtype=torch.int32).to(dev)
            self.per, self.lo, self.hi = shard_units(2 * nw, self.world, self.rank)
            self.n_mine = self.hi - self.lo
            # per-window conditioning is constant across steps: gather it once (all windows: every rank finishes the step)
            self.lq_t = torch.empty((nw, b, ch, T, T), dtype=torch.float32, device=dev)
            self.xc_t = torch.empty_like(self.lq_t)
            ops.tile_gather(lq, self.table_dev, T, self.lq_t)
            ops.tile_gather(xc, self.table_dev, T, self.xc_t)
            # conditioning rows of this rank's units (unit u: branch u // nw (0 = unconditional), window u % nw)
            def unit_cond(u, key):
                if u < nw:
                    return self.uc[key]
                return (cond[u - nw] if self.use_local_prompt else cond)[key]
            if self.n_mine > 0:
                units = range(self.lo, self.hi)
                self.u_ctx = torch.cat([unit_cond(u, "crossattn") for u in units], 0).contiguous()
                self.u_vec = torch.cat([unit_cond(u, "vector") for u in units], 0).contiguous()
                self.u_lq = torch.cat([self.lq_t[u % nw] for u in units], 0).contiguous()
            # network outputs of every unit; padded to equal slots per rank only when the split is uneven
            self.sizes = [shard_units(2 * nw, self.world, r) for r in range(self.world)]
            self.even = all(hi - lo == self.per for _, lo, hi in self.sizes)
            self.units_pad = torch.zeros((self.per * self.world, b, ch, T, T), dtype=torch.float32, device=dev)
            self.units_all = self.units_pad if self.even else torch.zeros((2 * nw, b, ch, T, T), dtype=torch.float32, device=dev)
            self.tiles_out = torch.empty((nw, b, ch, T, T), dtype=torch.float32, device=dev)
        else:
            self.per, self.lo, self.hi = shard_windows(nw, self.world, self.rank)
            slots = self.per * self.world
            table = torch.full((slots, 4), -1, dtype=torch.int32)
            table[:nw] = torch.tensor(windows, dtype=torch.int32)
            self.table_dev = table.to(dev)
            self.my_table = self.table_dev[self.lo:self.hi].contiguous()
            self.tiles_out = torch.zeros((slots, b, ch, T, T), dtype=torch.float32, device=dev)
            self.n_mine = self.hi - self.lo
            if self.n_mine > 0:
                self.lq_t = torch.empty((self.n_mine, b, ch, T, T), dtype=torch.float32, device=dev)
                self.xc_t = torch.empty_like(self.lq_t)
                ops.tile_gather(lq, self.my_table, T, self.lq_t)
                ops.tile_gather(xc, self.my_table, T, self.xc_t)

    @torch.no_grad()
    def step(self, i, eps_noise=None):
        smp, x = self.smp, self.x
        k = smp.step_constants(self.sigmas, i, self.control_scale, self.lin_cs, self.cs_start)
        if eps_noise is None:
            eps_noise = torch.randn_like(x)     # drawn every step, on every rank, like the reference (sampling.py:631)
        x_next = self._step_fused(k, eps_noise) if self.fused else self._step_generic(k, eps_noise)
        self.x = x_next
        return x_next

    def _step_fused(self, k, eps_noise):
        smp, x = self.smp, self.x
        b, ch, T = self.shape
        nw, n_mine = self.nw, self.n_mine
        den = self.denoiser.denoiser
        sq, idx = den.quantize_host(k["sigma_hat"])
        c_in = den.c_in_host(sq)
        x_t = torch.empty((nw, b, ch, T, T), dtype=torch.float32, device=x.device)
        ops.tile_gather(x, self.table_dev, T, x_t)
        e_t = None
        if k["gamma"] > 0:
            e_t = torch.empty_like(x_t)
            ops.tile_gather(eps_noise, self.table_dev, T, e_t)
        x_hat = torch.empty_like(x_t)
        net_in = torch.empty((2 * nw, b, ch, T, T), dtype=torch.float32, device=x.device)     # unit u <-> block u (both halves equal)
        ops.edm_pre(x_t, e_t, k["noise_mul"], c_in, x_hat, net_in)
        if n_mine > 0:
            slot0 = self.rank * self.per
            for g0, g1 in balanced_groups(n_mine, 2 * smp.tile_batch):
                g = g1 - g0
                cg = {"control": self.u_lq[g0 * b:g1 * b], "crossattn": self.u_ctx[g0 * b:g1 * b], "vector": self.u_vec[g0 * b:g1 * b]}
                t = torch.full((g * b,), idx, dtype=torch.long, device=x.device)
                out = self.denoiser.run_network(net_in[self.lo + g0:self.lo + g1].reshape(g * b, ch, T, T), t, cg, k["control_scale"],
                                                (self.uid, g0))       # the conditioning is constant over this run's steps
                self.units_pad[slot0 + g0:slot0 + g1] = out.view(g, b, ch, T, T)
        if self.world > 1:
            # the ONE exchange of a step: every rank receives the raw network outputs of all units (NCCL all-gather over NVLink)
            exchange_unit_outputs(self.units_pad, self.rank, self.per, self.group)
            if not self.even:
                for r, (_, lo, hi) in enumerate(self.sizes):
                    self.units_all[lo:hi].copy_(self.units_pad[r * self.per:r * self.per + (hi - lo)])
        ops.edm_post(x_hat, self.units_all, self.xc_t if k["use_restore"] else None, -sq, smp.guider.scale_host(k["sigma_hat"]),
                     k["restore_mul"], k["sigma_hat"], k["dt"], self.tiles_out)
        x_next = torch.empty_like(x)
        ops.tile_blend(self.tiles_out, self.table_dev, T, self.weights64, x_next)
        return x_next

    def _step_generic(self, k, eps_noise):
        smp, x = self.smp, self.x
        b, ch, T = self.shape
        n_mine, lo = self.n_mine, self.lo
        cond, uc = self.cond, self.uc
        if n_mine > 0:
            x_t = torch.empty((n_mine, b, ch, T, T), dtype=torch.float32, device=x.device)
            ops.tile_gather(x, self.my_table, T, x_t)
            e_t = None
            if k["gamma"] > 0:
                e_t = torch.empty_like(x_t)
                ops.tile_gather(eps_noise, self.my_table, T, e_t)
            for g0, g1 in balanced_groups(n_mine, smp.tile_batch):
                g = g1 - g0

                def stack(key):
                    return torch.cat([(cond[lo + j] if self.use_local_prompt else cond)[key] for j in range(g0, g1)], 0)
                c_g = {"control": self.lq_t[g0:g1].reshape(g * b, ch, T, T), "crossattn": stack("crossattn"),
                       "vector": stack("vector")}
                uc_g = {"control": c_g["control"], "crossattn": torch.cat([uc["crossattn"]] * g, 0),
                        "vector": torch.cat([uc["vector"]] * g, 0)}
                out = smp._step(self.denoiser, x_t[g0:g1].reshape(g * b, ch, T, T),
                                None if e_t is None else e_t[g0:g1].reshape(g * b, ch, T, T), c_g, uc_g,
                                self.xc_t[g0:g1].reshape(g * b, ch, T, T), k)
                self.tiles_out[lo + g0:lo + g1] = out.view(g, b, ch, T, T)
        if self.world > 1:
            exchange_window_outputs(self.tiles_out, self.rank, self.per, self.group)
        x_next = torch.empty_like(x)
        ops.tile_blend(self.tiles_out, self.table_dev, T, self.weights64, x_next)
        return x_next


# ----------------------------------------------------------------------------------------------------------------------
# DPM++ 2M SDE restore samplers (SUPIR_v0_Juggernautv9_lightning.yaml) — reference sampling.py:271-360, 422-515, 663-730
# ----------------------------------------------------------------------------------------------------------------------
def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """k-diffusion 0.1.1 `get_sigmas_karras` (the reference imports it, sampling.py:20; the package is not vendored):
    Karras et al. 2022 schedule with an appended zero, restated from the published formula."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = float(sigma_min) ** (1 / rho)
    max_inv_rho = float(sigma_max) ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])]).to(device)


class BrownianIncrementNoise:
    """Simplest stand-in for k-diffusion's BrownianTreeNoiseSampler: the reference asks it for the normalised Brownian
    increment between two sigmas, once per step, on disjoint intervals — i.e. independent N(0, I) draws; this class draws
    them with torch.randn_like. The samplers' default is supir_b200.brownian.BrownianTreeNoiseSampler (one fixed, seeded,
    query-order-independent Brownian path per run); neither reproduces torchsde's numbers (parity unpinned)."""

    def __init__(self, x, sigma_min=None, sigma_max=None):
        self.like = x

    def __call__(self, sigma, sigma_next):
        return torch.randn_like(self.like)


class RestoreDPMPP2MSampler(BaseDiffusionSampler):
    noise_sampler_cls = BrownianTreeNoiseSampler        # sampling.py:491-494, 684-687 (k-diffusion's class of the same name)

    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, restore_cfg=4.0, restore_cfg_s_tmin=0.05,
                 eta=1.0, *args, **kwargs):
        self.s_noise, self.eta = s_noise, eta
        super().__init__(*args, **kwargs)

    # host-side float32 step multipliers (get_variables / get_mult, sampling.py:290-315, 435-446)
    def step_multipliers(self, sigma, next_sigma, previous_sigma):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            t, t_next = f32(-np.log(f32(sigma))), f32(-np.log(f32(next_sigma)))
            h = f32(t_next - t)
            eta_h = f32(self.eta) * h
            m1 = f32(np.exp(-t_next) / np.exp(-t) * np.exp(-eta_h))
            m2 = f32(np.expm1(-h - eta_h))
            m3 = m4 = None
            if previous_sigma is not None:
                r = f32((t - f32(-np.log(f32(previous_sigma)))) / h)
                m3, m4 = f32(1 + 1 / (2 * r)), f32(1 / (2 * r))
            noise_mul = f32(next_sigma) * f32(np.sqrt(-np.expm1(-2 * eta_h))) * f32(self.s_noise)
        return float(m1), float(m2), (None if m3 is None else float(m3)), (None if m4 is None else float(m4)), float(noise_mul)

    def _denoise(self, denoiser, x, sigma, cond, uc, control_scale):
        N = x.shape[0]
        if isinstance(denoiser, FusedDenoiser) and hasattr(self.guider, "scale_host"):
            sq, idx = denoiser.denoiser.quantize_host(sigma)
            c_in = denoiser.denoiser.c_in_host(sq)
            x_hat = torch.empty_like(x)
            net_in = torch.empty((2 * N,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
            ops.edm_pre(x, None, 0.0, c_in, x_hat, net_in)
            cpair = {key: torch.cat((uc[key], cond[key]), 0) for key in ("vector", "crossattn", "control")}
            t = torch.full((2 * N,), idx, dtype=torch.long, device=x.device)
            net_out = denoiser.network(net_in, t, cpair, control_scale)
            den, scratch = torch.empty_like(x), torch.empty_like(x)
            ops.edm_post(x_hat, net_out, None, -sq, self.guider.scale_host(sigma), 0.0, 1.0, 0.0, scratch, denoised=den)
            return den
        st = torch.full((N,), sigma, dtype=torch.float32, device=x.device)
        xi, si, ci = self.guider.prepare_inputs(x, st, cond, uc)
        return self.guider(denoiser(xi, si, ci, control_scale), st).contiguous()

    def sampler_step(self, denoiser, old_denoised, previous_sigma, sigma, next_sigma, x, cond, uc, eps_noise, control_scale):
        denoised = self._denoise(denoiser, x, sigma, cond, uc, control_scale)
        m1, m2, m3, m4, noise_mul = self.step_multipliers(sigma, next_sigma, previous_sigma)
        out = torch.empty_like(x)
        if old_denoised is None or next_sigma < 1e-14:
            ops.axpby_f32(x, m1, denoised, -m2, out)
            return out, denoised
        dd = torch.empty_like(x)
        ops.axpby_f32(denoised, m3, old_denoised.contiguous(), -m4, dd)
        ops.axpby_f32(x, m1, dd, -m2, out)
        if self.eta:
            ops.axpby_f32(out, 1.0, eps_noise.contiguous(), noise_mul, dd)
            out = dd
        return out, denoised

    def _schedule(self, x, num_steps):
        sig = self.host_sigmas(num_steps)
        x0 = torch.empty_like(x, dtype=torch.float32)
        ops.axpby_f32(x.contiguous().float(), float(np.sqrt(f32(1.0) + sig[0] * sig[0])), None, 0.0, x0)
        # the reference asks k-diffusion for self.num_steps sigmas whatever `num_steps` overrides (sampling.py:474-476)
        return x0, get_sigmas_karras(self.num_steps, sig[-2], sig[0]).numpy().astype(np.float32), len(sig)

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, control_scale=1.0, **kwargs):
        run = self.begin(denoiser, x, cond, uc, num_steps, control_scale=control_scale)
        for i in range(run.num_steps):
            run.step(i)
        return run.x

    @torch.no_grad()
    def begin(self, denoiser, x, cond, uc=None, num_steps=None, control_scale=1.0, **kwargs):
        return _DPMRun(self, denoiser, x, cond, cond if uc is None else uc, num_steps, control_scale)


class _DPMRun:
    def __init__(self, smp, denoiser, x, cond, uc, num_steps, control_scale):
        self.smp, self.denoiser, self.cond, self.uc, self.control_scale = smp, denoiser, cond, uc, control_scale
        self.x, self.sigmas, num_sigmas = smp._schedule(x, num_steps)
        self.noise = smp.noise_sampler_cls(self.x, self.sigmas[-2], self.sigmas[0])
        self.old = None
        self.num_steps = num_sigmas - 1

    @torch.no_grad()
    def step(self, i):
        sig = self.sigmas
        eps = self.noise(sig[i], sig[i + 1]) if (i > 0 and sig[i + 1] > 1e-14) else None
        if i == 0:
            self.old = None
        self.x, self.old = self.smp.sampler_step(self.denoiser, self.old, None if i == 0 else float(sig[i - 1]), float(sig[i]),
                                                 float(sig[i + 1]), self.x, self.cond, self.uc, eps, self.control_scale)
        return self.x


class TiledRestoreDPMPP2MSampler(RestoreDPMPP2MSampler):
    def __init__(self, tile_size=128, tile_stride=64, tile_batch=8, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.tile_size, self.tile_stride, self.tile_batch = tile_size, tile_stride, max(1, int(tile_batch))
        self._weights = None

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, control_scale=1.0, **kwargs):
        use_local_prompt = isinstance(cond, list)
        b, ch, h, w = x.shape
        T = self.tile_size
        windows = _sliding_windows(h, w, T, self.tile_stride)
        nw = len(windows)
        lq = (cond[0] if use_local_prompt else cond)["control"].contiguous().float()
        uc = (cond[0] if use_local_prompt else cond) if uc is None else uc
        if self._weights is None:
            self._weights = gaussian_weights(T, T, 1, device=self.device)
        w64 = self._weights[0, 0].contiguous()
        table = torch.tensor(windows, dtype=torch.int32, device=x.device)
        x, sigmas, num_sigmas = self._schedule(x, num_steps)
        noise = self.noise_sampler_cls(x, sigmas[-2], sigmas[0])
        lq_t = torch.empty((nw, b, ch, T, T), dtype=torch.float32, device=x.device)
        ops.tile_gather(lq, table, T, lq_t)
        old = None
        for i in range(num_sigmas - 1):
            has_noise = i > 0 and sigmas[i + 1] > 1e-14
            eps_noise = noise(sigmas[i], sigmas[i + 1]) if has_noise else torch.zeros_like(x)
            x_t, e_t = torch.empty_like(lq_t), torch.empty_like(lq_t)
            ops.tile_gather(x, table, T, x_t)
            ops.tile_gather(eps_noise.contiguous(), table, T, e_t)
            o_t = None
            if old is not None:
                o_t = torch.empty_like(lq_t)
                ops.tile_gather(old, table, T, o_t)
            xs, ds = torch.empty_like(lq_t), torch.empty_like(lq_t)
            for g0, g1 in balanced_groups(nw, self.tile_batch):
                g = g1 - g0
                flat = lambda t_: t_[g0:g1].reshape(g * b, ch, T, T)  # noqa: E731
                c_g = {"control": flat(lq_t),
                       "crossattn": torch.cat([(cond[j] if use_local_prompt else cond)["crossattn"] for j in range(g0, g1)], 0),
                       "vector": torch.cat([(cond[j] if use_local_prompt else cond)["vector"] for j in range(g0, g1)], 0)}
                uc_g = {"control": c_g["control"], "crossattn": torch.cat([uc["crossattn"]] * g, 0), "vector": torch.cat([uc["vector"]] * g, 0)}
                _x, _d = self.sampler_step(denoiser, None if o_t is None else flat(o_t), None if i == 0 else float(sigmas[i - 1]),
                                           float(sigmas[i]), float(sigmas[i + 1]), flat(x_t), c_g, uc_g, flat(e_t), control_scale)
                xs[g0:g1] = _x.view(g, b, ch, T, T)
                ds[g0:g1] = _d.view(g, b, ch, T, T)
            x_next, old_next = torch.empty_like(x), torch.empty_like(x)
            ops.tile_blend(xs, table, T, w64, x_next)
            ops.tile_blend(ds, table, T, w64, old_next)
            x, old = x_next, old_next
        return x
