"""GPU parity of the VAE (untiled + tiled) and of the samplers against the oracle / the reference's golden outputs."""
import json
import os

import numpy as np
import pytest
import torch

from weights import make_state_dict, randn

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}


def rel_fro(a, b):
    return float((a - b).norm() / b.norm())


def test_vae_tiny_untiled_and_tiled_vs_reference():
    """bf16 storage vs the reference's fp32 run: relative Frobenius error <= 3e-2 (30 GroupNorm layers deep)."""
    from supir_b200 import vae
    g = np.load(os.path.join(G, "vae_tiny.npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=71)
    with torch.device("cuda"):
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=cfg, lossconfig={"target": "torch.nn.Identity"})
    ae.load_state_dict(sd, strict=True)
    img, z = randn((1, 3, 64, 48), 81).cuda(), randn((1, 4, 8, 6), 82).cuda()
    e = ae.encoder(img).cpu()
    d = ae.decoder(z).cpu()
    assert e.shape == g["enc_untiled"].shape and d.shape == g["dec_untiled"].shape
    r1, r2 = rel_fro(e, torch.from_numpy(g["enc_untiled"])), rel_fro(d, torch.from_numpy(g["dec_untiled"]))
    big, zbig = randn((1, 3, 192, 160), 83).cuda(), randn((1, 4, 40, 52), 84).cuda()
    he, hd = vae.VAEHook(ae.encoder, 64, is_decoder=False), vae.VAEHook(ae.decoder, 16, is_decoder=True)
    te, td = he(big).cpu(), hd(zbig).cpu()
    assert te.shape == g["enc_tiled"].shape and td.shape == g["dec_tiled"].shape
    r3, r4 = rel_fro(te, torch.from_numpy(g["enc_tiled"])), rel_fro(td, torch.from_numpy(g["dec_tiled"]))
    print(f"vae rel_fro: enc {r1:.4g} dec {r2:.4g} enc_tiled {r3:.4g} dec_tiled {r4:.4g}")
    assert max(r1, r2, r3, r4) <= 3e-2
    # small-image shortcut of VAEHook (tilevae.py:694-696)
    assert torch.equal(he(img).cpu(), e)
    # quant / post-quant + posterior
    from oracle import vae as ovae
    mom = ae.quant_conv(ae.encoder(img))
    ref_m = ovae.encode_moments(sd, img.cpu())
    assert rel_fro(mom.cpu(), ref_m) <= 3e-2
    dec = ae.decode(z).cpu()
    assert rel_fro(dec, ovae.decode(sd, z.cpu(), scale_factor=1.0)) <= 3e-2


def test_vae_fast_mode_vs_reference():
    """VAEHook fast mode (tilevae.py:776-817, 855-876: GroupNorm statistics estimated on a thumbnail, tiles independent) against
    the REFERENCE's fast-mode outputs, incl. the color_fix variant; same tolerance as the exact mode."""
    from supir_b200 import vae
    g = np.load(os.path.join(G, "vae_tiny.npz"))
    gf = np.load(os.path.join(G, "vae_tiny_fast.npz"))
    with torch.device("cuda"):
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=json.loads(str(g["cfg"])), lossconfig={"target": "torch.nn.Identity"})
    ae.load_state_dict(make_state_dict(json.loads(str(g["shapes"])), seed=71), strict=True)
    big, zbig = randn((1, 3, 192, 160), 83).cuda(), randn((1, 4, 40, 52), 84).cuda()
    he = vae.VAEHook(ae.encoder, 64, is_decoder=False, fast_encoder=True)
    hd = vae.VAEHook(ae.decoder, 16, is_decoder=True, fast_decoder=True)
    hc = vae.VAEHook(ae.encoder, 64, is_decoder=False, fast_encoder=True, color_fix=True)
    errs = [rel_fro(h(x).cpu(), torch.from_numpy(gf[k])) for h, x, k in ((he, big, "enc_tiled_fast"), (hd, zbig, "dec_tiled_fast"),
                                                                          (hc, big, "enc_tiled_fast_colorfix"))]
    print("vae fast mode rel_fro vs reference (enc, dec, enc+color_fix):", ["%.4g" % v for v in errs])
    assert max(errs) <= 3e-2
    std, mean = __import__("supir_b200.ops", fromlist=["ops"]).channel_std_mean(big)
    rs, rm = torch.std_mean(big, dim=[0, 2, 3], keepdim=True)
    assert torch.allclose(std, rs, rtol=1e-5, atol=1e-6) and torch.allclose(mean, rm, rtol=1e-5, atol=1e-6)


def test_vae_repacks_after_a_second_load_state_dict():
    """gradio_demo*.py switch checkpoints with model.load_state_dict(..., strict=False) at run time (gradio_demo_tiled.py:130,134):
    the kernel-layout weights must follow."""
    from supir_b200 import vae
    g = np.load(os.path.join(G, "vae_tiny.npz"))
    cfg = json.loads(str(g["cfg"]))
    shapes = json.loads(str(g["shapes"]))
    with torch.device("cuda"):
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=cfg, lossconfig={"target": "torch.nn.Identity"})
    ae.load_state_dict(make_state_dict(shapes, seed=71), strict=True)
    z = randn((1, 4, 8, 6), 82).cuda()
    d1 = ae.decoder(z).clone()
    ae.load_state_dict(make_state_dict(shapes, seed=72), strict=False)
    d2 = ae.decoder(z).clone()
    assert not torch.allclose(d1, d2)
    ae.load_state_dict(make_state_dict(shapes, seed=71), strict=False)
    assert torch.equal(ae.decoder(z), d1)


def toy_network(x, t, c, control_scale):
    tt = (t.float() / 1000.0).view(-1, 1, 1, 1)
    v = c["vector"].mean(dim=1).view(-1, 1, 1, 1)
    return 0.3 * torch.tanh(x) + 0.1 * tt + 0.2 * control_scale * c["control"] + 0.05 * v


class SeededNoise:
    def __init__(self, base):
        self.base, self.n = base, 0

    def __call__(self, x, **k):
        self.n += 1
        return randn(tuple(x.shape), self.base + self.n).to(x.device, x.dtype)


def make_denoiser():
    from supir_b200 import denoiser as dn
    return dn.DiscreteDenoiserWithControl(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config=DISC).cuda()


def test_samplers_with_toy_network_match_reference(monkeypatch):
    """Sampler/denoiser/guider arithmetic in fp32: 1e-4 against the reference's golden runs (same seeded noise)."""
    from supir_b200 import sampling
    g = np.load(os.path.join(G, "sampler_toy.npz"))
    den = make_denoiser()
    guider = {"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 1.0, "scale_min": 4.0}}
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    for fused in (False, True):
        denoiser = sampling.FusedDenoiser(den, toy_network) if fused else (lambda x, s, c, cs: den(toy_network, x, s, c, cs))
        for name, restore_cfg, lin_cs in [("edm", -1.0, False), ("edm_restore", 4.0, True)]:
            smp = sampling.RestoreEDMSampler(num_steps=6, restore_cfg=restore_cfg, s_churn=5, s_noise=1.01,
                                             discretization_config=DISC, guider_config=guider)
            x = randn((2, 4, 12, 10), 50)
            c = {"control": randn((2, 4, 12, 10), 51), "vector": randn((2, 6), 52), "crossattn": randn((2, 3, 5), 53)}
            uc = {"control": c["control"], "vector": randn((2, 6), 54), "crossattn": randn((2, 3, 5), 55)}
            xc = randn((2, 4, 12, 10), 56)
            monkeypatch.setattr(torch, "randn_like", SeededNoise(1000))
            out = smp(denoiser, x.cuda(), cond=cu(c), uc=cu(uc), x_center=xc.cuda(), control_scale=0.9,
                      use_linear_control_scale=lin_cs, control_scale_start=0.2).cpu()
            torch.testing.assert_close(out, torch.from_numpy(g[name]), rtol=1e-4, atol=1e-4)
        for tile_batch in (1, 3):
            smp = sampling.TiledRestoreEDMSampler(tile_size=16, tile_stride=8, tile_batch=tile_batch, num_steps=4, restore_cfg=4.0,
                                                  s_churn=5, s_noise=1.01, discretization_config=DISC, guider_config=guider)
            x = randn((1, 4, 40, 28), 60)
            c = {"control": randn((1, 4, 40, 28), 61), "vector": randn((1, 6), 62), "crossattn": randn((1, 3, 5), 63)}
            uc = {"control": c["control"], "vector": randn((1, 6), 64), "crossattn": randn((1, 3, 5), 65)}
            xc = randn((1, 4, 40, 28), 66)
            monkeypatch.setattr(torch, "randn_like", SeededNoise(2000))
            out = smp(denoiser, x.cuda(), cond=cu(c), uc=cu(uc), x_center=xc.cuda(), control_scale=1.0).cpu()
            torch.testing.assert_close(out, torch.from_numpy(g["tiled"]), rtol=1e-4, atol=1e-4)
            # per-window prompts (cond is a list, one dict per window)
            nwin = len(sampling._sliding_windows(40, 28, 16, 8))
            conds = [cu({"control": c["control"], "vector": randn((1, 6), 700 + j), "crossattn": randn((1, 3, 5), 800 + j)}) for j in range(nwin)]
            monkeypatch.setattr(torch, "randn_like", SeededNoise(2500))
            out = smp(denoiser, x.cuda(), cond=conds, uc=cu(uc), x_center=xc.cuda(), control_scale=1.0).cpu()
            torch.testing.assert_close(out, torch.from_numpy(g["tiled_local"]), rtol=1e-4, atol=1e-4)


def test_dpmpp_samplers_match_reference():
    """DPM++ 2M SDE restore samplers (Lightning config): 1e-4 against the reference's runs with the same injected noise."""
    from supir_b200 import sampling
    g = np.load(os.path.join(G, "sampler_dpmpp_toy.npz"))
    den = make_denoiser()
    guider = {"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 2.0, "scale_min": 2.0}}
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731

    class Brownian:
        def __init__(self, x, a=None, b=None):
            self.shape, self.n = tuple(x.shape), 0

        def __call__(self, a, b):
            self.n += 1
            return randn(self.shape, 3000 + self.n).cuda()

    for fused in (False, True):
        denoiser = sampling.FusedDenoiser(den, toy_network) if fused else (lambda x, s, c, cs: den(toy_network, x, s, c, cs))
        smp = sampling.RestoreDPMPP2MSampler(num_steps=5, s_noise=1.003, eta=1.0, discretization_config=DISC, guider_config=guider)
        smp.noise_sampler_cls = Brownian
        x = randn((1, 4, 12, 10), 70)
        c = {"control": randn((1, 4, 12, 10), 71), "vector": randn((1, 6), 72), "crossattn": randn((1, 3, 5), 73)}
        uc = {"control": c["control"], "vector": randn((1, 6), 74), "crossattn": randn((1, 3, 5), 75)}
        out = smp(denoiser, x.cuda(), cond=cu(c), uc=cu(uc), control_scale=0.9).cpu()
        torch.testing.assert_close(out, torch.from_numpy(g["dpmpp"]), rtol=1e-4, atol=1e-4)
        smp = sampling.TiledRestoreDPMPP2MSampler(tile_size=16, tile_stride=8, tile_batch=3, num_steps=4, s_noise=1.003, eta=1.0,
                                                  discretization_config=DISC, guider_config=guider)
        smp.noise_sampler_cls = Brownian
        x = randn((1, 4, 40, 28), 80)
        c = {"control": randn((1, 4, 40, 28), 81), "vector": randn((1, 6), 82), "crossattn": randn((1, 3, 5), 83)}
        uc = {"control": c["control"], "vector": randn((1, 6), 84), "crossattn": randn((1, 3, 5), 85)}
        out = smp(denoiser, x.cuda(), cond=cu(c), uc=cu(uc), control_scale=1.0).cpu()
        torch.testing.assert_close(out, torch.from_numpy(g["dpmpp_tiled"]), rtol=1e-4, atol=1e-4)


def test_brownian_tree_noise_on_the_device():
    """The DPM++ samplers' default noise source (supir_b200/brownian.py) with CUDA generators: reproducible from the seed,
    order-independent, unit variance, additive over adjacent intervals (statistics in depth: tests/test_brownian.py)."""
    from supir_b200.brownian import BrownianTreeNoiseSampler
    x = torch.zeros(1, 4, 256, 256, device="cuda")
    a, b = BrownianTreeNoiseSampler(x, 0.0292, 14.6146, seed=11), BrownianTreeNoiseSampler(x, 0.0292, 14.6146, seed=11)
    sig = [14.6146, 5.0, 1.2, 0.3, 0.0292]
    fwd = [a(sig[i], sig[i + 1]) for i in range(4)]
    bwd = [b(sig[i], sig[i + 1]) for i in reversed(range(4))][::-1]
    assert all(p.is_cuda and torch.equal(p, q) for p, q in zip(fwd, bwd))
    for z in fwd:
        assert abs(float(z.var()) - 1) < 0.03 and abs(float(z.mean())) < 0.02
    w = lambda s0, s1: a(s0, s1) * abs(s1 - s0) ** 0.5  # noqa: E731  un-normalised increment
    assert torch.allclose(w(5.0, 1.2) + w(1.2, 0.3), w(5.0, 0.3), atol=1e-5)
