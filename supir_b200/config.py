"""String-keyed factory compatible with the reference's plugin surface (sgm/util.py:168-185).

`target:` strings written for the reference (`sgm.modules...`, `SUPIR.modules...`) resolve to this package's classes when
they are on the sampling hot path; anything else is imported normally (so a conditioner from the real `sgm` package still
works if that package is installed).
"""
import importlib

_ALIASES = {
    "SUPIR.models.SUPIR_model.SUPIRModel": "supir_b200.model.SUPIRModel",
    "sgm.modules.diffusionmodules.wrappers.ControlWrapper": "supir_b200.wrappers.ControlWrapper",
    "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiserWithControl": "supir_b200.denoiser.DiscreteDenoiserWithControl",
    "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting": "supir_b200.denoiser.EpsWeighting",
    "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling": "supir_b200.denoiser.EpsScaling",
    "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization": "supir_b200.denoiser.LegacyDDPMDiscretization",
    "sgm.modules.diffusionmodules.guiders.LinearCFG": "supir_b200.guiders.LinearCFG",
    "sgm.modules.diffusionmodules.guiders.VanillaCFG": "supir_b200.guiders.VanillaCFG",
    "sgm.modules.diffusionmodules.guiders.IdentityGuider": "supir_b200.guiders.IdentityGuider",
    "sgm.modules.diffusionmodules.sampling_utils.NoDynamicThresholding": "supir_b200.guiders.NoDynamicThresholding",
    "sgm.modules.diffusionmodules.sampling.RestoreEDMSampler": "supir_b200.sampling.RestoreEDMSampler",
    "sgm.modules.diffusionmodules.sampling.TiledRestoreEDMSampler": "supir_b200.sampling.TiledRestoreEDMSampler",
    "sgm.modules.diffusionmodules.sampling.RestoreDPMPP2MSampler": "supir_b200.sampling.RestoreDPMPP2MSampler",
    "sgm.modules.diffusionmodules.sampling.TiledRestoreDPMPP2MSampler": "supir_b200.sampling.TiledRestoreDPMPP2MSampler",
    "SUPIR.modules.SUPIR_v0.GLVControl": "supir_b200.nets.GLVControl",
    "SUPIR.modules.SUPIR_v0.LightGLVUNet": "supir_b200.nets.LightGLVUNet",
    "sgm.models.autoencoder.AutoencoderKLInferenceWrapper": "supir_b200.vae.AutoencoderKLInferenceWrapper",
    "sgm.models.autoencoder.AutoencoderKL": "supir_b200.vae.AutoencoderKL",
    "sgm.modules.GeneralConditioner": "supir_b200.conditioner.GeneralConditioner",
    "sgm.modules.GeneralConditionerWithControl": "supir_b200.conditioner.GeneralConditionerWithControl",
    "sgm.modules.PreparedConditioner": "supir_b200.conditioner.PreparedConditioner",
    "sgm.modules.encoders.modules.PreparedConditioner": "supir_b200.conditioner.PreparedConditioner",
    "sgm.modules.encoders.modules.GeneralConditioner": "supir_b200.conditioner.GeneralConditioner",
    "sgm.modules.encoders.modules.GeneralConditionerWithControl": "supir_b200.conditioner.GeneralConditionerWithControl",
    "sgm.modules.encoders.modules.FrozenCLIPEmbedder": "supir_b200.conditioner.FrozenCLIPEmbedder",
    "sgm.modules.encoders.modules.FrozenOpenCLIPEmbedder2": "supir_b200.conditioner.FrozenOpenCLIPEmbedder2",
    "sgm.modules.encoders.modules.ConcatTimestepEmbedderND": "supir_b200.conditioner.ConcatTimestepEmbedderND",
}


def get_obj_from_str(string):
    string = _ALIASES.get(string, string)
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**dict(config.get("params", dict()) or {}))


class AttrDict(dict):
    """Nested-dict view with attribute access, enough of OmegaConf's DictConfig for what the reference's callers do with a
    loaded options/*.yaml (`config.model`, `config.SDXL_CKPT`, `default_setting.s_cfg_Quality`, and SUPIRModel's
    `sampler_config.params.num_steps = ...`)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value


def attrify(obj):
    if isinstance(obj, dict):
        return AttrDict({k: attrify(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [attrify(v) for v in obj]
    return obj


def load_yaml(path, attr_access=False):
    """options/*.yaml as plain dicts (OmegaConf is not required); attr_access=True wraps them in AttrDict."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    return attrify(cfg) if attr_access else cfg
