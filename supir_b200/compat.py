"""Drop-in installation into a checkout of the reference (Fanghua-Yu/SUPIR).

The reference resolves every hot-path class by dotted string through `sgm.util.instantiate_from_config`
(sgm/util.py:168-185) and `get_obj_from_str` (sgm/models/diffusion.py:50-53). `install()` imports the reference's own
modules and rebinds those names to this package's classes, so `test.py` / `gradio_demo_tiled.py` — which only touch the
object returned by `create_SUPIR_model` — run unchanged on the B200 backend:

    python -c "import supir_b200.compat as c; c.install(); import runpy; runpy.run_path('test.py', run_name='__main__')" ...

SUPIRModel itself (orchestration, text conditioner, colour fix, checkpoint loading) stays the reference's.
"""
import importlib

# reference module -> {attribute: replacement "module:attr" in this package}
PATCHES = {
    "sgm.modules.diffusionmodules.wrappers": {"ControlWrapper": "supir_b200.wrappers:ControlWrapper"},
    "sgm.modules.diffusionmodules.denoiser": {"DiscreteDenoiserWithControl": "supir_b200.denoiser:DiscreteDenoiserWithControl"},
    "sgm.modules.diffusionmodules.guiders": {"LinearCFG": "supir_b200.guiders:LinearCFG", "VanillaCFG": "supir_b200.guiders:VanillaCFG"},
    "sgm.modules.diffusionmodules.sampling": {
        "RestoreEDMSampler": "supir_b200.sampling:RestoreEDMSampler",
        "TiledRestoreEDMSampler": "supir_b200.sampling:TiledRestoreEDMSampler",
        "RestoreDPMPP2MSampler": "supir_b200.sampling:RestoreDPMPP2MSampler",
        "TiledRestoreDPMPP2MSampler": "supir_b200.sampling:TiledRestoreDPMPP2MSampler",
        "gaussian_weights": "supir_b200.sampling:gaussian_weights",
        "_sliding_windows": "supir_b200.sampling:_sliding_windows",
    },
    "SUPIR.modules.SUPIR_v0": {"GLVControl": "supir_b200.nets:GLVControl", "LightGLVUNet": "supir_b200.nets:LightGLVUNet",
                               "ZeroSFT": "supir_b200.nets:ZeroSFT", "ZeroCrossAttn": "supir_b200.nets:ZeroCrossAttn"},
    "sgm.models.autoencoder": {"AutoencoderKLInferenceWrapper": "supir_b200.vae:AutoencoderKLInferenceWrapper",
                               "AutoencoderKL": "supir_b200.vae:AutoencoderKL"},
    "SUPIR.utils.tilevae": {"VAEHook": "supir_b200.vae:VAEHook"},
    "SUPIR.models.SUPIR_model": {"VAEHook": "supir_b200.vae:VAEHook",
                                 "DiagonalGaussianDistribution": "supir_b200.vae:DiagonalGaussianDistribution",
                                 "wavelet_reconstruction": "supir_b200.colorfix:wavelet_reconstruction",
                                 "adaptive_instance_normalization": "supir_b200.colorfix:adaptive_instance_normalization"},
}


def _resolve(spec):
    mod, attr = spec.split(":")
    return getattr(importlib.import_module(mod), attr)


def install(strict=True):
    """Rebind the reference's hot-path names. Returns the list of (module, attribute) pairs that were patched."""
    done = []
    for modname, attrs in PATCHES.items():
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        for attr, spec in attrs.items():
            setattr(mod, attr, _resolve(spec))
            done.append((modname, attr))
    return done
