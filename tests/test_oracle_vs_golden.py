"""Pins the CPU oracle (oracle/) against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sampler as osamp
from oracle import unet as ounet
from oracle import vae as ovae
from weights import make_state_dict, randn

G = os.path.join(os.path.dirname(__file__), "golden")
torch.set_grad_enabled(False)


def bits(t):
    return np.asarray(t, dtype=np.float32).view(np.uint32).tolist()


@pytest.fixture(scope="module")
def book():
    with open(os.path.join(G, "bookkeeping.json")) as f:
        return json.load(f)


def test_sliding_windows_exact(book):
    for case in book["sliding_windows"]:
        got = [list(c) for c in osamp.sliding_windows(*case["args"])]
        assert got == case["windows"], case["args"]


def test_split_tiles_and_crop_exact(book):
    for case, crop in zip(book["split_tiles"], book["crop"]):
        h, w, tile, dec = case["args"]
        ib, ob = ovae.split_tiles(h, w, tile, dec)
        assert ib == case["in"] and ob == case["out"], case["args"]
        for i, o, ref in zip(ib, ob, crop["crops"]):
            th, tw = (i[3] - i[2]), (i[1] - i[0])
            th, tw = (th * 8, tw * 8) if dec else (th // 8, tw // 8)
            y0, y1, x0, x1 = ovae.crop_margins(th, tw, i, o, dec)
            idx = torch.arange(th * tw).view(th, tw)[y0:y1, x0:x1]
            assert [int(idx[0, 0]), int(idx[-1, -1]), idx.shape[0], idx.shape[1]] == ref


def test_best_tile_size_exact(book):
    for lb, ub, ref in book["best_tile"]:
        assert ovae.get_best_tile_size(lb, ub) == ref


def test_sigma_tables_bit_exact(book):
    for n, ref in book["sigmas"].items():
        assert bits(osamp.legacy_ddpm_sigmas(int(n)).numpy()) == ref, n
    assert bits(osamp.denoiser_sigma_table().numpy()) == book["denoiser_table"]
    probe = torch.tensor(np.array(book["sigma_to_idx"]["sigma"], dtype=np.uint32).view(np.float32))
    assert osamp.sigma_to_idx(osamp.denoiser_sigma_table(), probe).tolist() == book["sigma_to_idx"]["idx"]


def test_gaussian_weights_bit_exact():
    g = np.load(os.path.join(G, "gaussian_weights.npz"))
    w = osamp.gaussian_weights(128, 128)
    assert w.dtype == np.float64
    assert np.array_equal(np.tile(w, (1, 4, 1, 1)), g["w128"])
    assert np.array_equal(np.tile(osamp.gaussian_weights(16, 24), (2, 4, 1, 1)), g["w16x24"])


def test_glvcontrol_tiny():
    g = np.load(os.path.join(G, "glvcontrol_tiny.npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=11)
    x, xt = randn((2, 4, 16, 24), 1), randn((2, 4, 16, 24), 2)
    ctx, y = randn((2, 7, 48), 3), randn((2, 64), 4)
    hs = ounet.glv_control_forward(sd, x, torch.tensor([999, 401]), xt, ctx, y, cfg["model_channels"],
                                   cfg["num_head_channels"])
    assert len(hs) == 10
    for i, h in enumerate(hs):
        torch.testing.assert_close(h, torch.from_numpy(g[f"hs{i}"]), rtol=1e-4, atol=1e-4)


def test_zero_modules():
    g = np.load(os.path.join(G, "zero_modules.npz"))
    c, h, h_ori = randn((2, 32, 8, 12), 5), randn((2, 64, 8, 12), 6), randn((2, 64, 8, 12), 7)
    sd = {"m." + k: v for k, v in make_state_dict(json.loads(str(g["sft_shapes"])), seed=21).items()}
    torch.testing.assert_close(ounet.zero_sft(sd, "m", c, h, h_ori, 0.7), torch.from_numpy(g["sft_out"]), rtol=1e-4, atol=1e-4)
    sd = {"m." + k: v for k, v in make_state_dict(json.loads(str(g["sft_nc_shapes"])), seed=21).items()}
    torch.testing.assert_close(ounet.zero_sft(sd, "m", c, h, None, 0.4), torch.from_numpy(g["sft_nc_out"]), rtol=1e-4, atol=1e-4)
    sd = {"m." + k: v for k, v in make_state_dict(json.loads(str(g["zca_shapes"])), seed=21).items()}
    ctx, x = randn((2, 64, 8, 12), 8), randn((2, 128, 8, 12), 9)
    torch.testing.assert_close(ounet.zero_cross_attn(sd, "m", ctx, x, 0.9), torch.from_numpy(g["zca_out"]), rtol=1e-4, atol=1e-4)


def toy_network(x, t, c, control_scale):
    tt = (t.float() / 1000.0).view(-1, 1, 1, 1)
    v = c["vector"].mean(dim=1).view(-1, 1, 1, 1)
    return 0.3 * torch.tanh(x) + 0.1 * tt + 0.2 * control_scale * c["control"] + 0.05 * v


class SeededNoise:
    def __init__(self, base):
        self.base, self.n = base, 0

    def __call__(self, x):
        self.n += 1
        return randn(tuple(x.shape), self.base + self.n).to(x.dtype)


def test_samplers_toy_network():
    g = np.load(os.path.join(G, "sampler_toy.npz"))
    for name, restore_cfg, lin_cs in [("edm", -1.0, False), ("edm_restore", 4.0, True)]:
        smp = osamp.RestoreEDMSampler(num_steps=6, restore_cfg=restore_cfg, s_churn=5, s_noise=1.01, scale=1.0, scale_min=4.0,
                                      randn_like=SeededNoise(1000))
        x = randn((2, 4, 12, 10), 50)
        c = {"control": randn((2, 4, 12, 10), 51), "vector": randn((2, 6), 52), "crossattn": randn((2, 3, 5), 53)}
        uc = {"control": c["control"], "vector": randn((2, 6), 54), "crossattn": randn((2, 3, 5), 55)}
        xc = randn((2, 4, 12, 10), 56)
        out = smp(toy_network, x, c, uc, xc, control_scale=0.9, use_linear_control_scale=lin_cs, control_scale_start=0.2)
        torch.testing.assert_close(out, torch.from_numpy(g[name]), rtol=1e-5, atol=1e-5)
    smp = osamp.TiledRestoreEDMSampler(tile_size=16, tile_stride=8, num_steps=4, restore_cfg=4.0, s_churn=5, s_noise=1.01,
                                       scale=1.0, scale_min=4.0, randn_like=SeededNoise(2000))
    x = randn((1, 4, 40, 28), 60)
    c = {"control": randn((1, 4, 40, 28), 61), "vector": randn((1, 6), 62), "crossattn": randn((1, 3, 5), 63)}
    uc = {"control": c["control"], "vector": randn((1, 6), 64), "crossattn": randn((1, 3, 5), 65)}
    xc = randn((1, 4, 40, 28), 66)
    out = smp(toy_network, x, c, uc, xc, control_scale=1.0)
    torch.testing.assert_close(out, torch.from_numpy(g["tiled"]), rtol=1e-5, atol=1e-5)
    nwin = len(osamp.sliding_windows(40, 28, 16, 8))
    conds = [{"control": c["control"], "vector": randn((1, 6), 700 + j), "crossattn": randn((1, 3, 5), 800 + j)} for j in range(nwin)]
    smp.randn_like = SeededNoise(2500)
    out = smp(toy_network, x, conds, uc, xc, control_scale=1.0)
    torch.testing.assert_close(out, torch.from_numpy(g["tiled_local"]), rtol=1e-5, atol=1e-5)


def test_vae_tiny_untiled_and_tiled():
    g = np.load(os.path.join(G, "vae_tiny.npz"))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=71)
    img, z = randn((1, 3, 64, 48), 81), randn((1, 4, 8, 6), 82)
    torch.testing.assert_close(ovae.forward(sd, "encoder.", img, False), torch.from_numpy(g["enc_untiled"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ovae.forward(sd, "decoder.", z, True), torch.from_numpy(g["dec_untiled"]), rtol=1e-4, atol=1e-4)
    big, zbig = randn((1, 3, 192, 160), 83), randn((1, 4, 40, 52), 84)
    torch.testing.assert_close(ovae.tiled_forward(sd, "encoder.", big, 64, False), torch.from_numpy(g["enc_tiled"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ovae.tiled_forward(sd, "decoder.", zbig, 16, True), torch.from_numpy(g["dec_tiled"]), rtol=1e-4, atol=1e-4)


@pytest.mark.slow
def test_unet_fullwidth_depth1():
    path = os.path.join(G, "unet_fullwidth_depth1.npz")
    g = np.load(path)
    cfg = json.loads(str(g["cfg"]))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=31)
    x = randn((2, 4, 16, 16), 41)
    cond = {"control": randn((2, 4, 16, 16), 42), "crossattn": randn((2, 77, 2048), 43), "vector": randn((2, 2816), 44)}
    t = torch.tensor([950, 120])
    control = ounet.glv_control_forward(sd, cond["control"], t, x, cond["crossattn"], cond["vector"], cfg["model_channels"],
                                        cfg["num_head_channels"], prefix="control_model.")
    np.testing.assert_allclose([float(h.mean()) for h in control], g["control_mean"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose([float(h.std()) for h in control], g["control_std"], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(control[9], torch.from_numpy(g["control9"]), rtol=1e-3, atol=1e-3)
    out = ounet.light_glv_unet_forward(sd, x, t, cond["crossattn"], cond["vector"], control, 0.8, cfg["model_channels"],
                                       cfg["num_head_channels"], prefix="diffusion_model.")
    torch.testing.assert_close(out, torch.from_numpy(g["out"]), rtol=1e-3, atol=1e-3)


def test_colorfix_oracle():
    from oracle import colorfix as oc
    g = np.load(os.path.join(G, "colorfix.npz"))
    content, style = randn((1, 3, 70, 90), 900) * 0.5, randn((1, 3, 70, 90), 901) * 0.5 + 0.1
    hi, lo = oc.wavelet_decomposition(content)
    torch.testing.assert_close(hi, torch.from_numpy(g["high"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lo, torch.from_numpy(g["low"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(oc.wavelet_reconstruction(content, style), torch.from_numpy(g["wavelet"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(oc.adaptive_instance_normalization(content, style), torch.from_numpy(g["adain"]), rtol=1e-5, atol=1e-6)


class SeededBrownian:
    def __init__(self, shape):
        self.shape, self.n = shape, 0

    def __call__(self, a, b):
        self.n += 1
        return randn(self.shape, 3000 + self.n)


def test_dpmpp_samplers_toy_network():
    """Step arithmetic of the DPM++ restore samplers is pinned on the reference; the Karras schedule and the Brownian noise
    come from outside the reference (k-diffusion) and are supplied identically to both sides."""
    g = np.load(os.path.join(G, "sampler_dpmpp_toy.npz"))
    smp = osamp.RestoreDPMPP2MSampler(num_steps=5, s_noise=1.003, eta=1.0, scale=2.0, scale_min=2.0, noise_sampler=SeededBrownian((1, 4, 12, 10)))
    x = randn((1, 4, 12, 10), 70)
    c = {"control": randn((1, 4, 12, 10), 71), "vector": randn((1, 6), 72), "crossattn": randn((1, 3, 5), 73)}
    uc = {"control": c["control"], "vector": randn((1, 6), 74), "crossattn": randn((1, 3, 5), 75)}
    torch.testing.assert_close(smp(toy_network, x, c, uc, control_scale=0.9), torch.from_numpy(g["dpmpp"]), rtol=1e-5, atol=1e-5)
    smp = osamp.TiledRestoreDPMPP2MSampler(tile_size=16, tile_stride=8, num_steps=4, s_noise=1.003, eta=1.0, scale=2.0, scale_min=2.0,
                                           noise_sampler=SeededBrownian((1, 4, 40, 28)))
    x = randn((1, 4, 40, 28), 80)
    c = {"control": randn((1, 4, 40, 28), 81), "vector": randn((1, 6), 82), "crossattn": randn((1, 3, 5), 83)}
    uc = {"control": c["control"], "vector": randn((1, 6), 84), "crossattn": randn((1, 3, 5), 85)}
    torch.testing.assert_close(smp(toy_network, x, c, uc, control_scale=1.0), torch.from_numpy(g["dpmpp_tiled"]), rtol=1e-5, atol=1e-5)


def test_vae_fast_mode_oracle_vs_reference_golden():
    """VAEHook fast mode (GroupNorm statistics estimated on a thumbnail; tilevae.py:776-817, 855-876), incl. the color_fix
    variant that estimates only up to the first downsample."""
    g = np.load(os.path.join(G, "vae_tiny.npz"))
    gf = np.load(os.path.join(G, "vae_tiny_fast.npz"))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=71)
    big, zbig = randn((1, 3, 192, 160), 83), randn((1, 4, 40, 52), 84)
    for name, out in (("enc_tiled_fast", ovae.tiled_forward(sd, "encoder.", big, 64, False, fast=True)),
                      ("dec_tiled_fast", ovae.tiled_forward(sd, "decoder.", zbig, 16, True, fast=True)),
                      ("enc_tiled_fast_colorfix", ovae.tiled_forward(sd, "encoder.", big, 64, False, fast=True, color_fix=True))):
        ref = torch.from_numpy(gf[name])
        assert out.shape == ref.shape
        assert torch.allclose(out, ref, atol=2e-4, rtol=1e-3), (name, float((out - ref).abs().max()))
