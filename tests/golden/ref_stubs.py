"""Import the UNMODIFIED reference (Fanghua-Yu/SUPIR at /root/reference) in this container.

Only used by tests/golden/make_golden.py and tests/test_oracle_vs_reference.py (both skip when /root/reference is
absent, e.g. on the GPU box). The reference's third-party imports that are not installed here (omegaconf,
pytorch_lightning, k_diffusion, open_clip, kornia, xformers, diffusers) are replaced by inert stub modules BEFORE the
import; no reference source is edited or copied. CPU obstacles are monkey-patched as SURVEY.md §8c lists.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SUPIR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sgm"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch
    import torch.nn as nn

    if "omegaconf" not in sys.modules:
        class ListConfig(list):
            pass

        class DictConfig(dict):
            pass

        class OmegaConf:  # noqa: D401 - only the names the reference touches at import / ctor time
            @staticmethod
            def load(path):
                raise RuntimeError("OmegaConf stub: load() is not available")

        om = _stub("omegaconf", ListConfig=ListConfig, DictConfig=DictConfig, OmegaConf=OmegaConf)
        _stub("omegaconf.listconfig", ListConfig=ListConfig)
        om.listconfig = sys.modules["omegaconf.listconfig"]
    if "pytorch_lightning" not in sys.modules:
        def seed_everything(seed):
            import random
            import numpy as np
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            return seed

        _stub("pytorch_lightning", LightningModule=nn.Module, seed_everything=seed_everything)
    if "k_diffusion" not in sys.modules:
        def _na(*a, **k):
            raise RuntimeError("k_diffusion stub")

        kd = _stub("k_diffusion")
        kd.sampling = _stub("k_diffusion.sampling", get_sigmas_karras=_na, BrownianTreeNoiseSampler=_na)
    for name in ("open_clip", "kornia"):
        if name not in sys.modules:
            _stub(name)
    if "diffusers" not in sys.modules:
        d = _stub("diffusers")
        d.utils = _stub("diffusers.utils")
        d.utils.import_utils = _stub("diffusers.utils.import_utils", is_xformers_available=lambda: False)
    if "CKPT_PTH" not in sys.modules:
        _stub("CKPT_PTH", LLAVA_CLIP_PATH=None, LLAVA_MODEL_PATH=None, SDXL_CLIP1_PATH=None, SDXL_CLIP2_CKPT_PTH=None)


def import_reference():
    """Returns a namespace with the reference modules used by the golden generator."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        import torch
        from sgm.modules.diffusionmodules import sampling, denoiser, discretizer, guiders, wrappers, model as vae_model
        from sgm.modules.diffusionmodules import openaimodel
        from sgm.modules import attention
        from SUPIR.modules import SUPIR_v0
        import SUPIR.utils.devices as devices
        devices.device = torch.device("cpu")
        from SUPIR.utils import tilevae
        tilevae.xformer_attn_forward = tilevae.attn_forward  # same maths
        tilevae.is_xformers_available = True
    ns = types.SimpleNamespace(sampling=sampling, denoiser=denoiser, discretizer=discretizer, guiders=guiders,
                               wrappers=wrappers, vae_model=vae_model, openaimodel=openaimodel, attention=attention,
                               SUPIR_v0=SUPIR_v0, tilevae=tilevae, devices=devices)
    return ns
