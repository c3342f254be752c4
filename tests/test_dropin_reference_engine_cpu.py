"""The drop-in claim, end to end, in the build container: the REFERENCE's own engine class (SUPIR/models/SUPIR_model.py
`SUPIRModel(DiffusionEngine)`, imported unmodified from /root/reference) is built twice from the same YAML-shaped config and the
same weights — once as is (pure PyTorch reference), once after `supir_b200.compat.install(conditioner=True)` so that every
`target:` string resolves to this package's classes — and `batchify_sample(image, prompts, ...)` is run on both: stage-1
encode / decode, re-encode, text conditioning from prompt strings, the EDM restore sampler driving the reference's opaque denoiser
lambda (SUPIR_model.py:123-125), final decode, also with the tiled VAE hooks (`init_tile_vae`), the tiled sampler and per-window prompts.
No GPU here, so the backend's kernels are the plain-torch stand-ins of tests/cpu_ops.py (bf16 storage where the kernels store
bf16): what this pins is that the reference's orchestration code runs UNCHANGED on the backend's classes — constructor
signatures, attributes it pokes (`model.dtype`, `load_control_model`, re-bindable `.forward` / `.original_forward`,
`sampler_config.params.*`), call signatures, RNG order — and produces the reference's result within the bf16 tolerance."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_stubs  # noqa: E402
from weights import cond_tokens, make_state_dict, randn  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_stubs.reference_available(), reason="/root/reference not present (GPU box)")
HERE = os.path.dirname(os.path.abspath(__file__))
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
# full SDXL widths (LightGLVUNet's 'XL-base' adapter tables are hard-coded for 320 / 640 / 1280 channels, SUPIR_v0.py:560-598) with
# depth-1 transformers, and both text towers at their real widths with two blocks each: 980 M + 145 M parameters per engine
NET = dict(adm_in_channels=2816, num_classes="sequential", use_checkpoint=True, in_channels=4, out_channels=4, model_channels=320,
           attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4], num_head_channels=64, use_spatial_transformer=True,
           use_linear_in_transformer=True, transformer_depth=[1, 1, 1], context_dim=2048, spatial_transformer_attn_type="softmax-xformers",
           legacy=False)
VAE = dict(attn_type="vanilla", double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4, 4],
           num_res_blocks=1, attn_resolutions=[], dropout=0.0)
ARCH_L = dict(vocab=49408, width=768, heads=12, layers=2, mlp=3072, ctx=77, act="quick_gelu", eps=1e-5)
ARCH_G = dict(vocab=49408, width=1280, heads=20, layers=2, mlp=5120, ctx=77, proj=1280, act="gelu", eps=1e-5)
EMB = [
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenCLIPEmbedder",
     "params": {"layer": "hidden", "layer_idx": 1, "arch": {"layers": 2}}},
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenOpenCLIPEmbedder2",
     "params": {"arch": "ViT-bigG-14", "layer": "penultimate", "always_return_pooled": True, "legacy": False, "text_cfg": {"layers": 2}}}]
EMB += [{"is_trainable": False, "input_key": k, "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}}
        for k in ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple")]


class Cfg(dict):
    """OmegaConf stand-in (the package is absent): nested dicts with attribute access, which batchify_sample uses
    (`self.sampler_config.params.num_steps = ...`, SUPIR_model.py:101-111)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return cfgify({k: copy.deepcopy(v, memo) for k, v in self.items()})


def cfgify(d):
    if isinstance(d, dict):
        return Cfg({k: cfgify(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [cfgify(v) for v in d]
    return d


def engine_config(sampler_target, with_conditioner, extra=None):
    return cfgify(dict(
        control_stage_config={"target": "SUPIR.modules.SUPIR_v0.GLVControl", "params": dict(NET, input_upscale=1)},
        network_config={"target": "SUPIR.modules.SUPIR_v0.LightGLVUNet", "params": dict(NET, mode="XL-base", project_type="ZeroSFT", project_channel_scale=2)},
        network_wrapper="sgm.modules.diffusionmodules.wrappers.ControlWrapper",
        denoiser_config={"target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiserWithControl",
                         "params": {"num_idx": 1000, "weighting_config": {"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
                                    "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                                    "discretization_config": DISC}},
        first_stage_config={"target": "sgm.models.autoencoder.AutoencoderKLInferenceWrapper",
                            "params": {"embed_dim": 4, "ddconfig": VAE, "lossconfig": {"target": "torch.nn.Identity"}}},
        conditioner_config=({"target": "sgm.modules.GeneralConditionerWithControl", "params": {"emb_models": EMB}} if with_conditioner else None),
        sampler_config={"target": sampler_target,
                        "params": dict({"num_steps": 100, "restore_cfg": 4.0, "s_churn": 0, "s_noise": 1.003, "device": "cpu", "discretization_config": DISC,
                                        "guider_config": {"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 7.5, "scale_min": 4.0}}},
                                       **(extra or {}))},
        ae_dtype="bf16", diffusion_dtype="bf16", scale_factor=0.13025, p_p=", best quality", n_p="blurry"))


def prompt_row(text, pad):
    import zlib
    return cond_tokens(zlib.crc32(text.encode()) % 100000, 1, 49408, pad)[0]


class SeededNoise:
    def __init__(self, base):
        self.base, self.n = base, 0

    def __call__(self, x, **k):
        self.n += 1
        return randn(tuple(x.shape), self.base + self.n).to(x.dtype)


def rel_fro(a, b):
    return float((a - b).norm() / b.norm())


@pytest.fixture()
def backend_on_standins(monkeypatch):
    """compat.install(conditioner=True) + CPU stand-in kernels + the three `needs CUDA` guards lifted; everything is put back."""
    import importlib
    ref_stubs.import_reference()
    sys.path.insert(0, HERE)
    import cpu_ops
    import supir_b200.compat as compat
    from supir_b200 import ops, vae, wrappers
    saved = {}
    for modname, attrs in dict(compat.PATCHES, **compat.CONDITIONER_PATCHES).items():
        m = importlib.import_module(modname)
        saved.update({(m, a): getattr(m, a) for a in attrs if hasattr(m, a)})

    def activate():
        cpu_ops.install(monkeypatch)
        monkeypatch.setattr(vae._VAENet, "_check", lambda self, x: (None if getattr(self, "_packed", False) else self.pack()))

        def conv1x1(self, x, in_scale=1.0):
            x = x.float().contiguous()
            out = torch.empty((x.shape[0], self.out_channels) + tuple(x.shape[2:]), dtype=torch.float32)
            w = self.weight.detach().reshape(self.out_channels, self.in_channels).to(torch.bfloat16).float().contiguous()
            return ops.conv1x1_small_nchw(x, w, vae._bias_bf16_values(self.bias), out, in_scale=in_scale)
        monkeypatch.setattr(vae._Conv1x1Small, "forward", conv1x1)

        def wrapper_forward(self, x, t, c, control_scale=1, context_token=None, **kw):      # ControlWrapper.forward minus the CUDA graph
            if not self._packed:
                self.pack()
            ctx, y = c["crossattn"], c["vector"]
            plan = wrappers._Plan(self, x.shape[0], x.shape[2], x.shape[3], ctx.shape[1], ctx.shape[2], y.shape[1], "cpu")
            wrappers.ControlWrapper._fill(plan, x, t, ctx, y, c["control"], control_scale, context_token)
            plan._run()
            return plan.out.clone()
        monkeypatch.setattr(wrappers.ControlWrapper, "forward", wrapper_forward)
        return compat.install(strict=True, conditioner=True)
    yield activate
    for (m, a), v in saved.items():
        setattr(m, a, v)


def build_reference_engine(RS, cfg, tok_l, tok_g):
    """Pure reference: its conditioner constructors need vocabularies / weights from the network, so the engine is built with the
    empty conditioner and the real one (tests/golden/make_golden.py:build_reference_conditioner) is attached afterwards."""
    import make_golden as mg
    from sgm.modules.encoders import modules as E
    eng = RS.SUPIRModel(**cfg)
    eng.conditioner = mg.build_reference_conditioner(E, tok_l, tok_g, ARCH_L, ARCH_G, layer_idx=1)
    return eng.eval()


SAMPLING = "sgm.modules.diffusionmodules.sampling."
P_P, N_P = ", best quality", "blurry"
LOCAL = ["a cat", "a dog on grass", "sky", "a red brick wall"]          # 24 x 24 latent, tile 16 / stride 8 -> 4 windows
# (name, sampler target, extra sampler params, tiled VAE, image size, prompts): `tiled_local_prompts` passes one prompt per
# sampler window, the way gradio_demo_tiled.py does (p = [[...]], SUPIR_model.py:163-176 -> a list of conds, sampling.py:623-627)
SCENARIOS = [("untiled", "RestoreEDMSampler", None, False, 128, ["a photo of a cat"]),
             ("tiled_local_prompts", "TiledRestoreEDMSampler", {"tile_size": 16, "tile_stride": 8}, True, 192, [LOCAL])]
KW = dict(num_steps=2, restoration_scale=4.0, s_churn=5, s_noise=1.01, cfg_scale=4.0, seed=77, control_scale=0.9,
          use_linear_CFG=True, cfg_scale_start=1.0, use_linear_control_scale=True, control_scale_start=0.3)


def run_scenarios(eng, monkeypatch, untile_reference_hooks=False):
    """batchify_sample for every scenario on ONE engine (the sampler is rebuilt per call from `sampler_config`, which test.py /
    the demos also overwrite between calls), then the stage-1-only path of gradio_demo*.py through a second denoise encoder."""
    out = {}
    for name, target, extra, tiled_vae, size, prompts in SCENARIOS:
        eng.sampler_config = engine_config(SAMPLING + target, False, extra)["sampler_config"]
        if tiled_vae and not getattr(eng, "_tile_vae_on", False):
            eng.init_tile_vae(encoder_tile_size=64, decoder_tile_size=8)
            eng._tile_vae_on = True
            if untile_reference_hooks:
                for net in (eng.first_stage_model.denoise_encoder, eng.first_stage_model.encoder, eng.first_stage_model.decoder):
                    net.forward.to_gpu = False                  # VAEHook(to_gpu=True) moves the net to devices.get_optimal_device()
        img = (randn((1, 3, size, size), 500) * 0.5).clamp(-1, 1)
        monkeypatch.setattr(torch, "randn_like", SeededNoise(9000))
        out[name] = eng.batchify_sample(img, list(prompts), **KW)
    eng.first_stage_model.denoise_encoder_s1 = copy.deepcopy(eng.first_stage_model.denoise_encoder)
    out["stage1_denoise"] = eng.batchify_denoise(img, is_stage1=True)       # gradio_demo_tiled.py:54, 95 (tiled hooks in place)
    return out


def test_reference_engine_runs_unchanged_on_the_backend(backend_on_standins, monkeypatch):
    import contextlib
    import io
    import warnings
    ref_stubs.import_reference()
    from SUPIR.models import SUPIR_model as RS
    texts = ["a photo of a cat" + P_P] + [t + P_P for t in LOCAL] + [N_P]
    tok_l = {t: prompt_row(t, 49407) for t in texts}
    tok_g = {t: prompt_row(t, 0) for t in texts}
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731
    cfg0 = SAMPLING + "RestoreEDMSampler"

    # ---- 1. the reference, untouched ----
    import SUPIR.modules.SUPIR_v0 as V
    assert V.GLVControl.__module__ == "SUPIR.modules.SUPIR_v0"
    with quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orig_tensor = torch.tensor

        def cpu_tensor(*a, **k):                                # gaussian_weights hard-codes device='cuda' (sampling.py:750)
            k.pop("device", None)
            return orig_tensor(*a, **k)
        monkeypatch.setattr(torch, "tensor", cpu_tensor)
        ref = build_reference_engine(RS, engine_config(cfg0, False), tok_l, tok_g)
        shapes = {k: list(v.shape) for k, v in ref.state_dict().items()}
        sd = make_state_dict(shapes, seed=123)
        sd["denoiser.sigmas"] = ref.denoiser.sigmas.clone()     # a persistent buffer, not a weight: keep the real sigma table
        ref.load_state_dict(sd)
        want = run_scenarios(ref, monkeypatch, untile_reference_hooks=True)
        assert type(ref.sampler).__module__ == "sgm.modules.diffusionmodules.sampling"
    del ref

    # ---- 2. the same class, same config + conditioner config, after compat.install(conditioner=True) ----
    done = backend_on_standins()
    assert ("SUPIR.modules.SUPIR_v0", "LightGLVUNet") in done and ("sgm.modules", "GeneralConditionerWithControl") in done
    with quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        eng = RS.SUPIRModel(**engine_config(cfg0, True)).eval()
        assert type(eng.model).__module__ == "supir_b200.wrappers" and type(eng.conditioner).__module__ == "supir_b200.conditioner"
        assert type(eng.first_stage_model).__module__ == "supir_b200.vae" and type(eng.denoiser).__module__ == "supir_b200.denoiser"
        assert {k: list(v.shape) for k, v in eng.state_dict().items()} == shapes, "state_dict layout differs from the reference engine's"
        eng.load_state_dict(sd)
        eng.conditioner.embedders[0].tokenize = lambda ts: torch.stack([tok_l[t] for t in ts])
        eng.conditioner.embedders[1].tokenize = lambda ts: torch.stack([tok_g[t] for t in ts])
        got = run_scenarios(eng, monkeypatch)
        assert type(eng.sampler).__module__ == "supir_b200.sampling" and type(eng.first_stage_model.decoder.forward).__module__ == "supir_b200.vae"
    errs = {k: rel_fro(got[k], want[k]) for k in want}
    print("reference SUPIRModel on the backend, rel. Frobenius vs the pure reference:", {k: f"{v:.4g}" for k, v in errs.items()})
    for k, v in errs.items():
        assert got[k].shape == want[k].shape and v <= (3e-2 if k == "stage1_denoise" else 5e-2), (k, v)
