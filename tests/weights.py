"""Deterministic synthetic weights shared by the golden generator and the tests.

Fixtures never store weight tensors: they store `{key: shape}` and a seed, and every party (the reference in the build
container, the oracle, the CUDA backend) regenerates the same state_dict from them. Zero-initialised tensors of the
reference (zero_module) receive random values too, otherwise parity through them would be vacuous.
"""
import zlib

import torch


def make_state_dict(shapes, seed=0, dtype=torch.float32):
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        leaf = key.rsplit(".", 1)[-1]
        leaf = {"in_proj_weight": "weight", "in_proj_bias": "bias", "positional_embedding": "weight", "text_projection": "weight"}.get(leaf, leaf)
        if leaf == "weight" and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / fan_in ** 0.5)
        elif leaf == "weight":           # norm scales
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        else:                             # buffers we do not model (never on the hot path)
            t = torch.zeros(shape)
        sd[key] = t.to(dtype)
    return sd


def shapes_of(module_or_sd):
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return {k: list(v.shape) for k, v in sd.items()}


def randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


# ---- text-conditioner fixtures (tests/golden/conditioner.npz): reduced towers, synthetic token rows ----
COND_L = dict(vocab=1000, width=128, heads=2, layers=4, mlp=512, ctx=77, act="quick_gelu", eps=1e-5)     # CLIP-L tower, reduced
COND_G = dict(vocab=1000, width=192, heads=3, layers=4, mlp=768, ctx=77, proj=160, act="gelu", eps=1e-5)  # bigG tower, reduced
COND_LAYER_IDX = 3        # SUPIR_v0.yaml: layer_idx 11 of 12 = the input of the last block; here 3 of 4


def cond_tokens(seed, n, vocab, pad):
    """[SOT] words [EOT] pad...: EOT is the largest id, like in the CLIP vocabulary (the pooling looks for it with argmax)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.full((n, 77), pad, dtype=torch.long)
    for i in range(n):
        k = int(torch.randint(3, 40, (1,), generator=g))
        out[i, 0] = vocab - 2
        out[i, 1:1 + k] = torch.randint(1, vocab - 2, (k,), generator=g)
        out[i, 1 + k] = vocab - 1
    return out


def cond_batches():
    prompts, neg = ["a photo of a cat", "an oil painting"], "blurry"
    tl = {p: cond_tokens(100 + i, 1, 1000, 999)[0] for i, p in enumerate(prompts + [neg])}      # CLIPTokenizer pads with EOT
    tg = {p: cond_tokens(200 + i, 1, 1000, 0)[0] for i, p in enumerate(prompts + [neg])}        # open_clip.tokenize pads with 0
    control = randn((2, 4, 8, 8), 301)
    sizes = {"original_size_as_tuple": torch.tensor([[1024, 1024], [768, 1280]]), "crop_coords_top_left": torch.tensor([[0, 0], [16, 32]]),
             "target_size_as_tuple": torch.tensor([[1024, 1024], [1024, 1024]])}
    batch = dict(sizes, txt=list(prompts), control=control)
    batch_uc = dict(sizes, txt=[neg, neg], control=control)
    return batch, batch_uc, tl, tg
