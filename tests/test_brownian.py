"""The native Brownian-interval noise source of the DPM++ restore samplers (supir_b200/brownian.py): the
properties the reference relies on from k-diffusion's BrownianTreeNoiseSampler (absent here, see the module docstring) —
reproducible from the seed, independent of query order, additive over adjacent intervals, independent over disjoint ones,
Brownian covariance, unit-variance normalised increments, k-diffusion's sign convention."""
import math

import torch

from supir_b200.brownian import BrownianTree, BrownianTreeNoiseSampler

N = 400_000


def test_reproducible_and_order_independent():
    x = torch.zeros(2, 4, 8, 8)
    sig = [14.6, 9.1, 4.2, 1.3, 0.4, 0.03]
    a = BrownianTreeNoiseSampler(x, 0.03, 14.6, seed=123)
    b = BrownianTreeNoiseSampler(x, 0.03, 14.6, seed=123)
    fwd = [a(sig[i], sig[i + 1]) for i in range(5)]
    bwd = [b(sig[i], sig[i + 1]) for i in reversed(range(5))][::-1]
    assert all(torch.equal(p, q) for p, q in zip(fwd, bwd))
    c = BrownianTreeNoiseSampler(x, 0.03, 14.6, seed=124)
    assert not torch.equal(c(sig[0], sig[1]), fwd[0])
    assert torch.equal(a(sig[1], sig[0]), -fwd[0])                 # BatchedBrownianTree.sort: swapped arguments flip the sign
    torch.manual_seed(7)
    d = BrownianTreeNoiseSampler(x, 0.03, 14.6)
    torch.manual_seed(7)
    e = BrownianTreeNoiseSampler(x, 0.03, 14.6)
    assert d.seed == e.seed and torch.equal(d(3.0, 2.0), e(3.0, 2.0))     # seed=None: drawn from the global generator


def test_increments_add_up_and_match_brownian_moments():
    tree = BrownianTree(0.5, 10.5, (N,), entropy=99)
    pts = [0.5, 0.9, 2.0, 3.14159, 5.5, 5.5000001, 8.0, 10.5]
    for a, b, c in zip(pts, pts[1:], pts[2:]):
        assert torch.allclose(tree(a, b) + tree(b, c), tree(a, c), atol=1e-5)
    incs = [tree(a, b) for a, b in zip(pts, pts[1:])]
    for (a, b), w in zip(zip(pts, pts[1:]), incs):
        dt = b - a
        if dt < 1e-3:
            continue                                               # inside one leaf: linear interpolation, not a fresh increment
        assert abs(float(w.mean())) < 5 * math.sqrt(dt / N)
        assert abs(float(w.var()) / dt - 1) < 0.02, (a, b, float(w.var()) / dt)
    big = [w for (a, b), w in zip(zip(pts, pts[1:]), incs) if b - a > 1e-3]
    for i in range(len(big)):
        for j in range(i + 1, len(big)):
            corr = float((big[i] * big[j]).mean() / (big[i].std() * big[j].std()))
            assert abs(corr) < 5 / math.sqrt(N), (i, j, corr)
    ws, wt = tree.value(2.0), tree.value(8.0)                      # Cov(W(s), W(t)) = min(s, t) - t0
    assert abs(float((ws * wt).mean()) / 1.5 - 1) < 0.03
    assert float(tree.value(0.5).abs().max()) == 0.0


def test_normalised_sampler_has_unit_variance_and_normal_tails():
    x = torch.zeros(N)
    s = BrownianTreeNoiseSampler(x, 0.0292, 14.6146, seed=5)
    for a, b in ((14.6146, 6.0), (6.0, 1.7), (1.7, 0.2), (0.2, 0.0292)):
        z = s(a, b)
        assert abs(float(z.var()) - 1) < 0.02 and abs(float(z.mean())) < 0.01
        assert abs(float((z.abs() > 1.959964).float().mean()) - 0.05) < 0.004       # two-sided 5 % tail of a normal
        kurt = float((z ** 4).mean() / z.var() ** 2)
        assert abs(kurt - 3.0) < 0.1


def test_dpmpp_sampler_uses_the_brownian_tree_by_default():
    from supir_b200 import sampling
    assert sampling.RestoreDPMPP2MSampler.noise_sampler_cls is BrownianTreeNoiseSampler
    assert sampling.TiledRestoreDPMPP2MSampler.noise_sampler_cls is BrownianTreeNoiseSampler
