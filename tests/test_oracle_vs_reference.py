"""Live re-derivation of the committed golden fixtures from the UNMODIFIED reference, whenever /root/reference is present
(the build container; the GPU box has no reference and skips this file): the cheap generators of tests/golden/make_golden.py
are re-run into a scratch directory and must reproduce the committed files bit for bit. Together with
tests/test_oracle_vs_golden.py (oracle == fixtures) this pins oracle == reference without trusting a stale fixture."""
import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_stubs  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_stubs.reference_available(), reason="/root/reference not present (GPU box)")
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def regen(tmp_path_factory):
    mg = importlib.import_module("make_golden")
    out = tmp_path_factory.mktemp("golden_live")
    saved = mg.HERE
    mg.HERE = str(out)
    try:
        mg.gen_bookkeeping()
        mg.gen_zero_modules()
        mg.gen_glvcontrol_tiny()
        mg.gen_sampler()
        mg.gen_colorfix()
        mg.gen_conditioner()
        mg.gen_vae_fast()
    finally:
        mg.HERE = saved
    return str(out)


def _same_npz(a, b):
    fa, fb = np.load(a), np.load(b)
    assert sorted(fa.files) == sorted(fb.files)
    for k in fa.files:
        x, y = fa[k], fb[k]
        assert x.shape == y.shape and x.dtype == y.dtype, k
        assert x.tobytes() == y.tobytes(), f"{os.path.basename(a)}:{k} differs from the committed fixture"


def test_bookkeeping_and_weights_regenerate_bit_identically(regen):
    with open(os.path.join(regen, "bookkeeping.json")) as f, open(os.path.join(G, "bookkeeping.json")) as g:
        assert f.read() == g.read()
    _same_npz(os.path.join(regen, "gaussian_weights.npz"), os.path.join(G, "gaussian_weights.npz"))


@pytest.mark.parametrize("name", ["zero_modules.npz", "glvcontrol_tiny.npz", "sampler_toy.npz", "colorfix.npz", "conditioner.npz", "vae_tiny_fast.npz"])
def test_tensor_fixtures_regenerate_bit_identically(regen, name):
    _same_npz(os.path.join(regen, name), os.path.join(G, name))


def test_pil2tensor_matches_reference_function():
    """supir_b200.util.PIL2Tensor (host-side input preparation, restated) against the reference's own function
    (SUPIR/util.py:60-84, executed from the read-only checkout) on random image sizes (24 of them) and every argument combination
    test.py / gradio_demo use: same working size, same reported size, same pixels."""
    import random
    import torch
    pytest.importorskip("PIL")
    from PIL import Image
    from supir_b200 import util
    src = open("/root/reference/SUPIR/util.py").read()
    ns = {"np": np, "torch": torch, "Image": Image}
    exec(src[src.index("def PIL2Tensor"):src.index("def Tensor2PIL")], ns)
    rng = random.Random(0)
    for _ in range(24):
        w, h = rng.randint(40, 900), rng.randint(40, 900)
        img = Image.fromarray(np.random.RandomState(w * 1000 + h).randint(0, 255, (h, w, 3), dtype=np.uint8))
        for up, ms, fr in [(1, 1024, None), (2, 1024, None), (4, 64, None), (1, 1024, 512), (3, 256, 512)]:
            a = ns["PIL2Tensor"](img, upsacle=up, min_size=ms, fix_resize=fr)
            b = util.PIL2Tensor(img, upsacle=up, min_size=ms, fix_resize=fr)
            assert a[1:] == b[1:] and torch.equal(a[0], b[0]), (w, h, up, ms, fr)
