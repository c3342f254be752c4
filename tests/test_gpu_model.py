"""End-to-end GPU test of the engine (supir_b200.model.SUPIRModel.batchify_sample: stage-1 encode/decode, re-encode with the
CPU-generator posterior sample, untiled RestoreEDMSampler with the fused step kernels, final decode) against the same
pipeline composed from the CPU oracle with identical weights, inputs and noise. (The full-depth SDXL configuration of the
control + UNet pair is checked in tests/test_gpu_bench_networks.py.)"""
import json
import os

import numpy as np
import pytest
import torch

from weights import make_state_dict, randn

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}


class SeededNoise:
    def __init__(self, base):
        self.base, self.n = base, 0

    def __call__(self, x, **k):
        self.n += 1
        return randn(tuple(x.shape), self.base + self.n).to(x.device, x.dtype)


def rel_fro(a, b):
    return float((a - b).norm() / b.norm())


# the conditioner of options/SUPIR_v0.yaml:66-105 at the REAL widths (768 + 1280 = 2048 context channels, 1280 + 3 * 512 = 2816
# vector channels) with two transformer blocks per tower, so that the CPU oracle stays cheap
COND_CFG = {"target": "sgm.modules.GeneralConditionerWithControl", "params": {"emb_models": [
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenCLIPEmbedder",
     "params": {"layer": "hidden", "layer_idx": 1, "arch": {"layers": 2}}},
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenOpenCLIPEmbedder2",
     "params": {"arch": "ViT-bigG-14", "version": "laion2b_s39b_b160k", "freeze": True, "layer": "penultimate", "always_return_pooled": True,
                "legacy": False, "text_cfg": {"layers": 2}}},
    {"is_trainable": False, "input_key": "original_size_as_tuple", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}},
    {"is_trainable": False, "input_key": "crop_coords_top_left", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}},
    {"is_trainable": False, "input_key": "target_size_as_tuple", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}}]}}


def prompt_tokens(text, pad):
    """Deterministic stand-in for the two BPE tokenisers (no vocabulary files offline): [SOT] ids(text) [EOT] pad..."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(text.encode()))
    k = 3 + len(text) % 40
    row = torch.full((77,), pad, dtype=torch.long)
    row[0], row[1:1 + k], row[1 + k] = 49406, torch.randint(1, 49405, (k,), generator=g), 49407
    return row


def attach_fake_tokenizers(conditioner):
    conditioner.embedders[0].tokenize = lambda texts: torch.stack([prompt_tokens(t, 49407) for t in texts])   # CLIPTokenizer pads with EOT
    conditioner.embedders[1].tokenize = lambda texts: torch.stack([prompt_tokens(t, 0) for t in texts])       # open_clip pads with 0


def oracle_condition(sd_cond, z, prompts):
    from oracle import textenc as otext
    N = len(prompts)
    batch = {"original_size_as_tuple": torch.tensor([1024, 1024]).repeat(N, 1), "crop_coords_top_left": torch.tensor([0, 0]).repeat(N, 1),
             "target_size_as_tuple": torch.tensor([1024, 1024]).repeat(N, 1), "control": z,
             "txt_tokens_l": torch.stack([prompt_tokens(t, 49407) for t in prompts]), "txt_tokens_g": torch.stack([prompt_tokens(t, 0) for t in prompts])}
    return otext.supir_conditioner(sd_cond, batch, 12, 20, clip_layer_idx=1)


@pytest.mark.parametrize("with_prompts", [False, True])
def test_engine_batchify_sample_vs_oracle_pipeline(monkeypatch, with_prompts):
    """with_prompts: the conditioning comes from prompt strings through the engine's own kernel-backed text conditioner
    (prepare_condition, SUPIR_model.py:152-179) instead of ready-made c / uc dictionaries."""
    from oracle import sampler as osamp, unet as ounet, vae as ovae
    from supir_b200 import model as smodel
    gu = np.load(os.path.join(G, "unet_fullwidth_depth1.npz"))
    gv = np.load(os.path.join(G, "vae_tiny.npz"))
    ucfg, vcfg = json.loads(str(gu["cfg"])), json.loads(str(gv["cfg"]))
    sd_net = make_state_dict(json.loads(str(gu["shapes"])), seed=31)
    sd_vae = make_state_dict(json.loads(str(gv["shapes"])), seed=71)
    sd_vae.update({"denoise_encoder." + k[len("encoder."):]: v for k, v in sd_vae.items() if k.startswith("encoder.")})
    cfg = dict(
        control_stage_config={"target": "SUPIR.modules.SUPIR_v0.GLVControl", "params": dict(ucfg, input_upscale=1)},
        network_config={"target": "SUPIR.modules.SUPIR_v0.LightGLVUNet",
                        "params": dict(ucfg, mode="XL-base", project_type="ZeroSFT", project_channel_scale=2)},
        network_wrapper="sgm.modules.diffusionmodules.wrappers.ControlWrapper",
        denoiser_config={"target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiserWithControl",
                         "params": {"num_idx": 1000, "weighting_config": {"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
                                    "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                                    "discretization_config": DISC}},
        first_stage_config={"target": "sgm.models.autoencoder.AutoencoderKLInferenceWrapper",
                            "params": {"embed_dim": 4, "ddconfig": vcfg, "lossconfig": {"target": "torch.nn.Identity"}}},
        sampler_config={"target": "sgm.modules.diffusionmodules.sampling.RestoreEDMSampler",
                        "params": {"num_steps": 100, "restore_cfg": 4.0, "s_churn": 0, "s_noise": 1.003, "discretization_config": DISC,
                                   "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 7.5, "scale_min": 4.0}}}},
        ae_dtype="bf16", diffusion_dtype="bf16", scale_factor=0.13025, p_p=", masterpiece", n_p="blurry, lowres")
    if with_prompts:
        cfg["conditioner_config"] = COND_CFG
    with torch.device("cuda"):
        m = smodel.SUPIRModel(**cfg)
    m.model.load_state_dict(sd_net)
    m.first_stage_model.load_state_dict(sd_vae)
    if with_prompts:
        assert m.conditioner is not None, getattr(m, "_conditioner_error", None)
        sd_cond = make_state_dict({k: list(v.shape) for k, v in m.conditioner.state_dict().items()}, seed=77)
        m.conditioner.load_state_dict(sd_cond)
        attach_fake_tokenizers(m.conditioner)
    img = (randn((1, 3, 128, 128), 200) * 0.5).clamp(-1, 1)
    c = {"crossattn": randn((1, 77, 2048), 201), "vector": randn((1, 2816), 202)}
    uc = {"crossattn": randn((1, 77, 2048), 203), "vector": randn((1, 2816), 204)}
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    steps, seed = 3, 1234
    monkeypatch.setattr(torch, "randn_like", SeededNoise(5000))
    cond_args = dict(p=["a photo of a cat"]) if with_prompts else dict(c=cu(c), uc=cu(uc))
    out = m.batchify_sample(img.cuda(), num_steps=steps, restoration_scale=4.0, s_churn=5, s_noise=1.01, cfg_scale=4.0, seed=seed,
                            control_scale=0.9, use_linear_CFG=True, cfg_scale_start=1.0, **cond_args).cpu()
    # ---- the same pipeline from the oracle ----
    noise = SeededNoise(5000)
    torch.manual_seed(seed)
    _z = ovae.gaussian_latent(ovae.encode_moments(sd_vae, img, encoder_prefix="denoise_encoder."), None)
    x_stage1 = ovae.decode(sd_vae, _z)
    mom = ovae.encode_moments(sd_vae, x_stage1)
    z_stage1 = ovae.gaussian_latent(mom, torch.randn(mom.shape[0], 4, *mom.shape[2:]))
    smp = osamp.RestoreEDMSampler(num_steps=steps, restore_cfg=4.0, s_churn=5, s_noise=1.01, scale=1.0, scale_min=4.0, randn_like=noise)
    net = lambda x, t, cc, cs: ounet.control_wrapper_forward(sd_net, x, t, cc, cs)  # noqa: E731
    noised = noise(_z)
    if with_prompts:
        c = oracle_condition(sd_cond, _z, ["a photo of a cat, masterpiece"])
        uc = oracle_condition(sd_cond, _z, ["blurry, lowres"])
    zs = smp(net, noised, dict(c, control=_z), dict(uc, control=_z), z_stage1, control_scale=0.9)
    ref = ovae.decode(sd_vae, zs)
    e = rel_fro(out, ref)
    print(f"engine end-to-end ({'prompts -> image' if with_prompts else 'c / uc given'}) rel_fro={e:.4g}")
    assert out.shape == ref.shape and e <= 5e-2
