"""Shared by tests/test_conditioner_cpu.py and tests/test_gpu_conditioner.py: the product's conditioner built exactly as
options/SUPIR_v0.yaml:66-105 spells it (reduced towers, tests/golden/conditioner.npz) and its comparison with the reference's
golden outputs."""
import json
import os

import numpy as np
import torch

from weights import COND_G, COND_L, COND_LAYER_IDX, make_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "conditioner.npz"))


def rel_fro(a, b):
    return float((a - b).norm() / b.norm())


def golden_sd():
    return make_state_dict(json.loads(str(G["shapes"])), seed=91)


EMB_MODELS = [
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenCLIPEmbedder",
     "params": {"layer": "hidden", "layer_idx": COND_LAYER_IDX, "arch": COND_L}},
    {"is_trainable": False, "input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenOpenCLIPEmbedder2",
     "params": {"arch": "ViT-bigG-14", "version": "laion2b_s39b_b160k", "freeze": True, "layer": "penultimate", "always_return_pooled": True,
                "legacy": False, "text_cfg": COND_G}},
    {"is_trainable": False, "input_key": "original_size_as_tuple", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}},
    {"is_trainable": False, "input_key": "crop_coords_top_left", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}},
    {"is_trainable": False, "input_key": "target_size_as_tuple", "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}},
]


def build_product_conditioner(tl, tg, device="cpu"):
    """The conditioner exactly as options/SUPIR_v0.yaml:66-105 spells it (reduced towers), through the string-keyed factory."""
    from supir_b200.config import instantiate_from_config
    gc = instantiate_from_config({"target": "sgm.modules.GeneralConditionerWithControl", "params": {"emb_models": EMB_MODELS}})
    shapes = json.loads(str(G["shapes"]))
    assert {k: list(v.shape) for k, v in gc.state_dict().items()} == shapes, "state_dict layout differs from the reference's"
    gc.load_state_dict(make_state_dict(shapes, seed=91), strict=True)
    gc = gc.to(device)
    gc.embedders[0].tokenize = lambda texts: torch.stack([tl[t] for t in texts])
    gc.embedders[1].tokenize = lambda texts: torch.stack([tg[t] for t in texts])
    return gc


def check_against_golden(gc, batch, batch_uc, tol):
    worst = 0.0
    c, uc = gc.get_unconditional_conditioning(dict(batch), dict(batch_uc))
    _, uc0 = gc.get_unconditional_conditioning(dict(batch), dict(batch_uc), force_uc_zero_embeddings=["txt"])
    for name, d in (("c", c), ("uc", uc), ("uc0", uc0)):
        assert torch.equal(d["control"], batch["control"])           # the control latent rides along untouched (modules.py:241)
        for k in ("crossattn", "vector"):
            ref = torch.from_numpy(G[f"{name}_{k}"])
            got = d[k].float().cpu()
            assert got.shape == ref.shape and got.dtype == torch.float32
            if float(ref.norm()) == 0.0:                      # force_zero_embeddings: exact zeros, not "small"
                assert float(got.abs().max()) == 0.0, (name, k)
                continue
            e = rel_fro(got, ref)
            worst = max(worst, e)
            assert e <= tol, (name, k, e)
    return worst


