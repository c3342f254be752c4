"""Reproducible Brownian-interval noise for the DPM++ 2M SDE restore samplers — a native replacement for
k-diffusion's `BrownianTreeNoiseSampler` (k_diffusion/sampling.py, pinned 0.1.1.post1; a wrapper over `torchsde.BrownianTree`),
which the reference builds once per run and queries once per step (sgm/modules/diffusionmodules/sampling.py:491-494, 684-687).
Neither package is under /root/reference or in this image, so the torchsde stream itself cannot be reproduced; what the sampler
needs from it is restated here:

  * ONE Brownian path W on [t_min, t_max] per run; the noise of a step is the increment W(t_b) - W(t_a) of that path, divided
    by sqrt(|t_b - t_a|) (unit variance), with the sign convention of k-diffusion's BatchedBrownianTree (swapped arguments flip
    the sign);
  * increments over disjoint intervals are independent, increments over adjacent intervals add up (the property a fixed-path
    SDE solver relies on and that independent draws per call do not have when intervals are re-queried or subdivided);
  * the path is a pure function of (seed, t): values do not depend on the order or history of queries, and a run is
    reproducible from the seed. With seed=None the seed is drawn from torch's global generator, like k-diffusion does, so
    `torch.manual_seed` fixes the whole run (and all ranks of a sharded run draw the same path).

Construction: W(t_min) = 0, W(t_max) = sqrt(T) z_root, then dyadic Brownian-bridge refinement — the midpoint of an interval
[l, r] is (W(l) + W(r)) / 2 + sqrt((r - l) / 4) z_node — down to `depth` levels and linear interpolation inside a leaf (the
legacy torchsde BrownianTree does the same below its tolerance). Each node's z comes from its own torch.Generator seeded with a
hash of (seed, level, index): that is what makes the path order-independent. PyTorch owns the RNG (plumbing): there is no kernel
here, and the arithmetic is a handful of axpy's on a latent-sized tensor once per step.
"""
import hashlib
import math

import torch


def _node_seed(entropy, level, index):
    h = hashlib.blake2b(f"{int(entropy)}:{int(level)}:{int(index)}".encode(), digest_size=8).digest()
    return int.from_bytes(h, "little") & ((1 << 63) - 1)


class BrownianTree:
    """W(t) for t in [t0, t1] as tensors of `shape`; W(t0) = 0."""

    def __init__(self, t0, t1, shape, device="cpu", dtype=torch.float32, entropy=0, depth=24):
        assert t1 > t0, (t0, t1)
        self.t0, self.t1, self.shape, self.device, self.dtype = float(t0), float(t1), tuple(shape), torch.device(device), dtype
        self.entropy, self.depth = int(entropy), int(depth)
        self._cache = {}                                  # (level, index) -> value at that node: shared upper path of the queries

    def _normal(self, level, index):
        g = torch.Generator(device=self.device)
        g.manual_seed(_node_seed(self.entropy, level, index))
        return torch.randn(self.shape, generator=g, device=self.device, dtype=self.dtype)

    def value(self, t):
        t = float(t)
        assert self.t0 - 1e-9 <= t <= self.t1 + 1e-9, (t, self.t0, self.t1)
        t = min(max(t, self.t0), self.t1)
        lo, hi = self.t0, self.t1
        wlo = torch.zeros(self.shape, device=self.device, dtype=self.dtype)
        whi = self._cache.get((0, 0))
        if whi is None:
            whi = self._cache[(0, 0)] = math.sqrt(hi - lo) * self._normal(0, 0)
        if t == hi:
            return whi
        index = 0
        for level in range(1, self.depth + 1):
            mid = 0.5 * (lo + hi)
            wmid = self._cache.get((level, index))
            if wmid is None:
                wmid = 0.5 * (wlo + whi) + math.sqrt((hi - lo) / 4.0) * self._normal(level, index)
                if level <= 4:                           # keep the top of the tree (every query crosses it): at most 31 tensors
                    self._cache[(level, index)] = wmid
            if t == mid:
                return wmid
            if t < mid:
                hi, whi, index = mid, wmid, 2 * index
            else:
                lo, wlo, index = mid, wmid, 2 * index + 1
        return wlo + (whi - wlo) * ((t - lo) / (hi - lo))

    def __call__(self, ta, tb):
        return self.value(tb) - self.value(ta)


class BrownianTreeNoiseSampler:
    """Same constructor and call as k-diffusion's class: sampler(sigma, sigma_next) -> N(0, I)-distributed tensor like x, the
    normalised increment of one fixed Brownian path between transform(sigma) and transform(sigma_next)."""

    def __init__(self, x, sigma_min, sigma_max, seed=None, transform=lambda x: x):
        self.transform = transform
        t0, t1 = float(transform(torch.as_tensor(float(sigma_min)))), float(transform(torch.as_tensor(float(sigma_max))))
        self.sign = 1.0
        if t0 > t1:
            t0, t1, self.sign = t1, t0, -1.0
        if seed is None:
            seed = int(torch.randint(0, 2 ** 63 - 1, []).item())
        self.seed = int(seed)
        self.tree = BrownianTree(t0, t1, x.shape, x.device, torch.float32, self.seed)

    def __call__(self, sigma, sigma_next):
        t0, t1 = float(self.transform(torch.as_tensor(float(sigma)))), float(self.transform(torch.as_tensor(float(sigma_next))))
        sign = self.sign
        if t0 > t1:
            t0, t1, sign = t1, t0, -sign
        return self.tree(t0, t1) * (sign / math.sqrt(abs(t1 - t0)))
