"""Text conditioner on the CPU: (1) the oracle against the reference's golden outputs and against the live
third-party towers that are installed (transformers' CLIPTextModel; open_clip's block restated on torch.nn.MultiheadAttention via
HF's gelu variant with mapped weights); (2) the PRODUCT's conditioner classes — state_dict layout, weight packing (stacked
Q|K|V, transposed text_projection), layer selection, pooling, concatenation order, unconditional branch — with the kernels
replaced by plain-torch stand-ins (tests/cpu_ops.py), against the same golden outputs. The kernels' arithmetic is pinned by
tests/test_gpu_conditioner.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import textenc as otext
from conditioner_util import G, build_product_conditioner, check_against_golden, golden_sd, rel_fro
from weights import COND_G, COND_L, COND_LAYER_IDX, cond_batches, make_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def oracle_batch(batch, tl, tg):
    return dict(batch, txt_tokens_l=torch.stack([tl[t] for t in batch["txt"]]), txt_tokens_g=torch.stack([tg[t] for t in batch["txt"]]))


def test_oracle_conditioner_vs_reference_golden():
    batch, batch_uc, tl, tg = cond_batches()
    sd = golden_sd()
    for name, b, zero in (("c", batch, ()), ("uc", batch_uc, ()), ("uc0", batch_uc, ("txt",))):
        out = otext.supir_conditioner(sd, oracle_batch(b, tl, tg), COND_L["heads"], COND_G["heads"], COND_LAYER_IDX, zero_keys=zero)
        for k in ("crossattn", "vector"):
            ref = torch.from_numpy(G[f"{name}_{k}"])
            assert out[k].shape == ref.shape
            assert torch.allclose(out[k], ref, atol=2e-5, rtol=1e-4), (name, k, float((out[k] - ref).abs().max()))


def test_oracle_clip_tower_vs_installed_transformers_full_config():
    """oracle.hf_clip_text_model against transformers.CLIPTextModel at the FULL openai/clip-vit-large-patch14 text config
    (random weights): every hidden state, the final layer norm and the pooled row."""
    tr = pytest.importorskip("transformers")
    cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                            max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)
    torch.manual_seed(0)
    m = tr.CLIPTextModel(cfg).eval()
    sd = {"transformer." + k: v for k, v in m.state_dict().items()}
    tokens = torch.randint(1, 49405, (2, 77))
    tokens[:, 0], tokens[0, 20:], tokens[1, 51:] = 49406, 49407, 49407
    ref = m(input_ids=tokens, output_hidden_states=True)
    hidden, last, pooled = otext.hf_clip_text_model(sd, tokens, 12)
    assert len(hidden) == len(ref.hidden_states) == 13
    for a, b in zip(hidden, ref.hidden_states):
        assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), float((a - b).abs().max())
    assert torch.allclose(last, ref.last_hidden_state, atol=2e-4, rtol=1e-4)
    assert torch.allclose(pooled, ref.pooler_output, atol=2e-4, rtol=1e-4)


def test_oracle_open_clip_tower_vs_hf_gelu_variant():
    """The open_clip text tower restated by the oracle against an INDEPENDENT implementation of the same published model:
    transformers.CLIPTextModelWithProjection with hidden_act='gelu' (how HF hosts laion/CLIP-ViT-bigG-14), weights mapped
    in_proj -> q/k/v, text_projection transposed."""
    tr = pytest.importorskip("transformers")
    a = COND_G
    sd = {k[len("embedders.1."):]: v for k, v in golden_sd().items() if k.startswith("embedders.1.")}
    cfg = tr.CLIPTextConfig(vocab_size=a["vocab"], hidden_size=a["width"], intermediate_size=a["mlp"], num_hidden_layers=a["layers"],
                            num_attention_heads=a["heads"], max_position_embeddings=a["ctx"], hidden_act="gelu", eos_token_id=2,
                            bos_token_id=0, pad_token_id=1, projection_dim=a["proj"])
    m = tr.CLIPTextModelWithProjection(cfg).eval()
    hf = {"text_model.embeddings.token_embedding.weight": sd["model.token_embedding.weight"],
          "text_model.embeddings.position_embedding.weight": sd["model.positional_embedding"],
          "text_model.final_layer_norm.weight": sd["model.ln_final.weight"], "text_model.final_layer_norm.bias": sd["model.ln_final.bias"],
          "text_projection.weight": sd["model.text_projection"].t().contiguous()}
    for i in range(a["layers"]):
        s, d = f"model.transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        for j, n in enumerate("qkv"):
            hf[d + f"self_attn.{n}_proj.weight"] = sd[s + "attn.in_proj_weight"].chunk(3, 0)[j]
            hf[d + f"self_attn.{n}_proj.bias"] = sd[s + "attn.in_proj_bias"].chunk(3, 0)[j]
        for src, dst in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"),
                         ("mlp.c_proj", "mlp.fc2")):
            hf[d + dst + ".weight"], hf[d + dst + ".bias"] = sd[s + src + ".weight"], sd[s + src + ".bias"]
    missing, unexpected = m.load_state_dict(hf, strict=False)
    assert not [k for k in missing if "position_ids" not in k] and not unexpected
    _, _, _, tg = cond_batches()
    tokens = torch.stack(list(tg.values()))
    ref = m(input_ids=tokens, output_hidden_states=True)
    o = otext.open_clip_text(sd, tokens, a["heads"])
    assert torch.allclose(o["penultimate"], ref.hidden_states[-2], atol=2e-5, rtol=1e-4)
    assert torch.allclose(o["last"], ref.hidden_states[-1], atol=2e-5, rtol=1e-4)
    assert torch.allclose(o["pooled"], ref.text_embeds, atol=2e-5, rtol=1e-4)


# ---- the product's classes on kernel stand-ins -----------------------------------------------------------------------------
@pytest.fixture()
def cpu_kernels(monkeypatch):
    sys.path.insert(0, HERE)
    import cpu_ops
    cpu_ops.install(monkeypatch)
    return cpu_ops


def test_product_conditioner_on_standins_vs_reference_golden(cpu_kernels):
    batch, batch_uc, tl, tg = cond_batches()
    gc = build_product_conditioner(tl, tg)
    worst = check_against_golden(gc, batch, batch_uc, tol=1.5e-2)     # bf16 operands in the stand-ins, fp32 reference
    print(f"conditioner on stand-ins vs reference golden: worst rel. Frobenius {worst:.3g}")


def test_product_embedder_layer_options_on_standins(cpu_kernels):
    """The other `layer` settings of both embedders against the oracle (last / pooled / hidden+pooled; legacy open_clip)."""
    from supir_b200 import conditioner as C
    _, _, tl, tg = cond_batches()
    sd = golden_sd()
    tok_l, tok_g = torch.stack(list(tl.values())), torch.stack(list(tg.values()))
    sd_l = {k[len("embedders.0."):]: v for k, v in sd.items() if k.startswith("embedders.0.")}
    sd_g = {k[len("embedders.1."):]: v for k, v in sd.items() if k.startswith("embedders.1.")}
    for layer, idx, pooled in (("last", None, False), ("pooled", None, False), ("hidden", -2, True), ("hidden", 0, False)):
        e = C.FrozenCLIPEmbedder(layer=layer, layer_idx=idx, always_return_pooled=pooled, arch=COND_L)
        e.load_state_dict(sd_l)
        got = e(tok_l)
        ref = otext.frozen_clip_embedder(sd_l, tok_l, COND_L["heads"], layer, idx, pooled)
        for g_, r_ in zip(got if pooled else [got], ref if pooled else [ref]):
            assert g_.shape == r_.shape and rel_fro(g_, r_) <= 1.5e-2, (layer, idx, rel_fro(g_, r_))
    for layer, legacy, pooled in (("last", True, False), ("penultimate", True, False), ("last", False, True), ("penultimate", False, False)):
        e = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", layer=layer, legacy=legacy, always_return_pooled=pooled, text_cfg=COND_G)
        e.load_state_dict(sd_g)
        got = e(tok_g)
        ref = otext.frozen_openclip_embedder2(sd_g, tok_g, COND_G["heads"], layer, pooled, legacy)
        for g_, r_ in zip(got if pooled else [got], ref if pooled else [ref]):
            assert g_.shape == r_.shape and rel_fro(g_, r_) <= 1.5e-2, (layer, legacy, rel_fro(g_, r_))


def test_repack_after_load_state_dict(cpu_kernels):
    """A second load_state_dict after a forward must change the output (weights are packed lazily and invalidated by the hook)."""
    from supir_b200 import conditioner as C
    _, _, tl, _ = cond_batches()
    tok = torch.stack(list(tl.values()))
    e = C.FrozenCLIPEmbedder(layer="last", arch=COND_L)
    shapes = {k: list(v.shape) for k, v in e.state_dict().items()}
    e.load_state_dict(make_state_dict(shapes, seed=1))
    a = e(tok).clone()
    e.load_state_dict(make_state_dict(shapes, seed=2))
    b = e(tok)
    assert rel_fro(a, b) > 0.1


def test_engine_prepare_condition_uses_the_conditioner(cpu_kernels):
    """SUPIRModel.prepare_condition (SUPIR_model.py:152-179) with the kernel-backed conditioner: positive prompt + p_p suffix,
    negative prompt, local (per-window) prompts."""
    from supir_b200 import model as M
    batch, _, tl, tg = cond_batches()
    eng = M.SUPIRModel.__new__(M.SUPIRModel)
    torch.nn.Module.__init__(eng)
    eng.conditioner = build_product_conditioner(tl, tg)
    seen = []
    for emb, table in ((eng.conditioner.embedders[0], tl), (eng.conditioner.embedders[1], tg)):
        def tok(texts, table=table):
            seen.append(list(texts))
            return torch.stack([table[t.replace(", best quality", "")] for t in texts])
        emb.tokenize = tok
    z = batch["control"]
    c, uc = eng.prepare_condition(z, ["a photo of a cat", "an oil painting"], ", best quality", "blurry", 2)
    assert seen[0] == ["a photo of a cat, best quality", "an oil painting, best quality"] and seen[2] == ["blurry", "blurry"]
    assert c["crossattn"].shape == (2, 77, COND_L["width"] + COND_G["width"]) and c["vector"].shape == (2, COND_G["proj"] + 3 * 512)
    assert c["control"] is z and torch.equal(uc["control"], z)           # the reference deep-copies the batch for uc
    local = ["a photo of a cat", "an oil painting", "blurry"]
    cl, ucl = eng.prepare_condition(z[:1], [local], "", "blurry", 1)
    assert isinstance(cl, list) and len(cl) == 3 and ucl["crossattn"].shape[0] == 1
    # the batched pass over all window prompts == the reference's one-call-per-window loop (SUPIR_model.py:163-176)
    for i, t in enumerate(local):
        ci, uci = eng.prepare_condition(z[:1], [t], "", "blurry", 1)
        assert torch.equal(cl[i]["crossattn"], ci["crossattn"]) and torch.equal(cl[i]["vector"], ci["vector"]) and cl[i]["control"] is z[:1] or torch.equal(cl[i]["control"], z[:1])
        assert torch.equal(ucl["crossattn"], uci["crossattn"]) and torch.equal(ucl["vector"], uci["vector"])


def test_constructors_pick_up_local_pretrained_files(tmp_path):
    """FrozenCLIPEmbedder(version=<HF checkpoint dir>) and FrozenOpenCLIPEmbedder2(version=<open_clip .bin>) load what the
    reference's constructors would load from the same paths (modules.py:462-463, 530-536)."""
    tr = pytest.importorskip("transformers")
    from supir_b200 import conditioner as C
    a = COND_L
    cfg = tr.CLIPTextConfig(vocab_size=a["vocab"], hidden_size=a["width"], intermediate_size=a["mlp"], num_hidden_layers=a["layers"],
                            num_attention_heads=a["heads"], max_position_embeddings=a["ctx"], hidden_act="quick_gelu", eos_token_id=2,
                            bos_token_id=0, pad_token_id=1)
    m = tr.CLIPTextModel(cfg)
    m.save_pretrained(tmp_path / "clip")
    e = C.FrozenCLIPEmbedder(version=str(tmp_path / "clip"), layer="hidden", layer_idx=2)
    assert e.arch["width"] == a["width"] and e.arch["layers"] == a["layers"]
    for k, v in m.state_dict().items():
        assert torch.equal(e.transformer.state_dict()[k], v), k
    g = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", text_cfg=COND_G)
    sd = {k: torch.randn_like(v) for k, v in g.model.state_dict().items()}
    torch.save(dict(sd, **{"visual.conv1.weight": torch.zeros(3)}), tmp_path / "open_clip_pytorch_model.bin")
    g2 = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", version=str(tmp_path / "open_clip_pytorch_model.bin"), text_cfg=COND_G)
    for k, v in sd.items():
        assert torch.equal(g2.model.state_dict()[k], v), k


def test_tokenisation_layouts_with_a_stand_in_bpe():
    """The two tokenisers' LAYOUT rules around the (absent) BPE vocabulary: HF CLIPTokenizer call arguments and EOT padding for
    CLIP-L (modules.py:485-494); open_clip.tokenize = [SOT] + ids + [EOT], truncation keeps EOT last, ZERO padding (modules.py:554)."""
    from supir_b200 import conditioner as C

    class FakeBPE:                                   # one id per word; what transformers.CLIPTokenizer would be asked for
        def __call__(self, text, add_special_tokens=True, truncation=False, max_length=None, padding=None, return_tensors=None, **kw):
            if isinstance(text, str):
                assert add_special_tokens is False
                return {"input_ids": [1000 + len(w) for w in text.split()]}
            assert truncation and max_length == 77 and padding == "max_length" and return_tensors == "pt"
            rows = []
            for t in text:
                ids = [C.SOT_TOKEN] + [1000 + len(w) for w in t.split()][:75] + [C.EOT_TOKEN]
                rows.append(ids + [C.EOT_TOKEN] * (77 - len(ids)))
            return {"input_ids": torch.tensor(rows)}

    l = C.FrozenCLIPEmbedder(layer="hidden", layer_idx=1, arch=dict(COND_L, layers=2))
    g = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", layer="penultimate", legacy=False, always_return_pooled=True, text_cfg=dict(COND_G, layers=2))
    l._tokenizer = g._tokenizer = FakeBPE()
    try:
        tl = l.tokenize(["a cat", "word " * 100])
        assert tl.shape == (2, 77) and tl[0].tolist()[:5] == [C.SOT_TOKEN, 1001, 1003, C.EOT_TOKEN, C.EOT_TOKEN] and int(tl[1, -1]) == C.EOT_TOKEN
        tg = g.tokenize(["a cat", "word " * 100])
        assert tg.shape == (2, 77) and tg[0].tolist()[:6] == [C.SOT_TOKEN, 1001, 1003, C.EOT_TOKEN, 0, 0]
        assert int(tg[1, 0]) == C.SOT_TOKEN and int(tg[1, -1]) == C.EOT_TOKEN and int((tg[1] == 0).sum()) == 0
        assert tg.argmax(-1).tolist() == [3, 76]     # the pooling finds EOT as the largest id
    finally:
        l._tokenizer = g._tokenizer = None


def test_towers_are_causal_and_row_independent(cpu_kernels):
    """Size-independent properties of the text towers: a token only influences its own and later positions (causal mask), and
    a prompt's embedding does not depend on what else is in the batch."""
    from supir_b200 import conditioner as C
    _, _, tl, tg = cond_batches()
    sd = golden_sd()
    for emb, toks, prefix in ((C.FrozenCLIPEmbedder(layer="hidden", layer_idx=COND_LAYER_IDX, arch=COND_L), tl, "embedders.0."),
                              (C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", layer="penultimate", legacy=False, text_cfg=COND_G), tg, "embedders.1.")):
        emb.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
        tok = torch.stack(list(toks.values()))
        base = emb(tok)
        changed = tok.clone()
        changed[:, 10] = (changed[:, 10] + 17) % 900 + 1
        out = emb(changed)
        assert torch.equal(out[:, :10], base[:, :10]) and not torch.equal(out[:, 10:], base[:, 10:])
        assert torch.equal(emb(tok[1:2]), base[1:2])


def test_ucg_rate_semantics(cpu_kernels):
    """modules.py:155-165, 181-192: an embedder with ucg_rate 1 is always dropped (zeros) in forward(), while
    get_unconditional_conditioning() switches the dropout off for its two passes and restores the rates afterwards."""
    batch, batch_uc, tl, tg = cond_batches()
    gc = build_product_conditioner(tl, tg)
    full = gc(dict(batch))
    gc.embedders[0].ucg_rate = 1.0
    dropped = gc(dict(batch))
    wl = COND_L["width"]
    assert float(dropped["crossattn"][..., :wl].abs().max()) == 0.0 and torch.equal(dropped["crossattn"][..., wl:], full["crossattn"][..., wl:])
    assert torch.equal(dropped["vector"], full["vector"])
    c, uc = gc.get_unconditional_conditioning(dict(batch), dict(batch_uc))
    assert torch.equal(c["crossattn"], full["crossattn"]) and gc.embedders[0].ucg_rate == 1.0
