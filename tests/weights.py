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
