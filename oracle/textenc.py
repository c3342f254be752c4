"""Oracle: the text conditioner — fp32, functional, driven by the reference's state_dict keys.
Test infrastructure only (see oracle/__init__.py).

The reference reaches the two text towers through third-party packages that are NOT under /root/reference:
  * transformers==4.28.1 `CLIPTextModel` (sgm/modules/encoders/modules.py:462-463, 494-503) — an installed transformers is in
    this image, so `tests/test_oracle_vs_reference.py` pins `hf_clip_text_model` against the live class;
  * open-clip-torch==2.17.1 `open_clip.create_model_and_transforms` (modules.py:530-536) — absent. Its text tower is restated
    from the published architecture (pre-LN ResidualAttentionBlock: x + attn(ln_1(x)); x + c_proj(gelu(c_fc(ln_2(x)))), built on
    torch.nn.MultiheadAttention with the causal `attn_mask`), and the reference's OWN glue code around it
    (`encode_with_transformer`, `text_transformer_forward`, `pool`, modules.py:567-607) is what the golden fixture
    tests/golden/conditioner.npz comes from: make_golden.py runs those unmodified methods over such blocks.
"""
import torch
import torch.nn.functional as F

from .unet import timestep_embedding


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _causal_attention(q, k, v, heads):
    """softmax(q k^T / sqrt(d) + causal mask) v over [B, L, W] (HF CLIPAttention; nn.MultiheadAttention with attn_mask)."""
    B, L, W = q.shape
    d = W // heads
    sp = lambda t: t.reshape(B, L, heads, d).transpose(1, 2)  # noqa: E731
    s = sp(q) @ sp(k).transpose(-1, -2) * d ** -0.5
    s = s + torch.full((L, L), float("-inf")).triu_(1)
    return (torch.softmax(s, dim=-1) @ sp(v)).transpose(1, 2).reshape(B, L, W)


def _act(x, kind):
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    return F.gelu(x)


# --------------------------------------------------------------------------------------------------------------------
# transformers.CLIPTextModel (behind FrozenCLIPEmbedder, modules.py:445-507)
# --------------------------------------------------------------------------------------------------------------------
def hf_clip_text_model(sd, tokens, heads, act="quick_gelu", prefix="transformer.text_model."):
    """Returns (hidden_states [n_layers + 1 entries], last_hidden_state, pooler_output) like
    CLIPTextModel(input_ids=tokens, output_hidden_states=True). pooler_output is taken at argmax(tokens) (the behaviour for
    eos_token_id == 2, which is what openai/clip-vit-large-patch14's config holds)."""
    p = prefix
    L = tokens.shape[1]
    x = sd[p + "embeddings.token_embedding.weight"][tokens] + sd[p + "embeddings.position_embedding.weight"][:L][None]
    hidden = [x]
    i = 0
    while f"{p}encoder.layers.{i}.layer_norm1.weight" in sd:
        lp = f"{p}encoder.layers.{i}."
        h = _ln(sd, lp + "layer_norm1", x)
        q, k, v = (F.linear(h, sd[lp + f"self_attn.{n}_proj.weight"], sd[lp + f"self_attn.{n}_proj.bias"]) for n in "qkv")
        a = _causal_attention(q, k, v, heads)
        x = x + F.linear(a, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        h = _ln(sd, lp + "layer_norm2", x)
        h = _act(F.linear(h, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]), act)
        x = x + F.linear(h, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])
        hidden.append(x)
        i += 1
    last = _ln(sd, p + "final_layer_norm", x)
    pooled = last[torch.arange(last.shape[0]), tokens.argmax(dim=-1)]
    return hidden, last, pooled


def frozen_clip_embedder(sd, tokens, heads, layer="last", layer_idx=None, return_pooled=False, act="quick_gelu", prefix=""):
    """FrozenCLIPEmbedder.forward after tokenisation (modules.py:494-507)."""
    hidden, last, pooled = hf_clip_text_model(sd, tokens, heads, act, prefix + "transformer.text_model.")
    z = last if layer == "last" else (pooled[:, None, :] if layer == "pooled" else hidden[layer_idx])
    return (z, pooled) if return_pooled else z


# --------------------------------------------------------------------------------------------------------------------
# open_clip text tower (behind FrozenOpenCLIPEmbedder2, modules.py:510-609)
# --------------------------------------------------------------------------------------------------------------------
def open_clip_text(sd, tokens, heads, prefix="model."):
    """encode_with_transformer with legacy=False (modules.py:567-607): {'penultimate', 'last'} residual streams (no ln_final)
    and 'pooled' = ln_final(last)[arange, tokens.argmax(-1)] @ text_projection."""
    p = prefix
    x = sd[p + "token_embedding.weight"][tokens] + sd[p + "positional_embedding"][None, :tokens.shape[1]]
    n = 0
    while f"{p}transformer.resblocks.{n}.ln_1.weight" in sd:
        n += 1
    out = {}
    for i in range(n):
        if i == n - 1:
            out["penultimate"] = x
        bp = f"{p}transformer.resblocks.{i}."
        h = _ln(sd, bp + "ln_1", x)
        q, k, v = F.linear(h, sd[bp + "attn.in_proj_weight"], sd[bp + "attn.in_proj_bias"]).chunk(3, dim=-1)
        a = _causal_attention(q, k, v, heads)
        x = x + F.linear(a, sd[bp + "attn.out_proj.weight"], sd[bp + "attn.out_proj.bias"])
        h = _ln(sd, bp + "ln_2", x)
        h = F.gelu(F.linear(h, sd[bp + "mlp.c_fc.weight"], sd[bp + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[bp + "mlp.c_proj.weight"], sd[bp + "mlp.c_proj.bias"])
    out["last"] = x
    o = _ln(sd, p + "ln_final", x)
    out["pooled"] = o[torch.arange(o.shape[0]), tokens.argmax(dim=-1)] @ sd[p + "text_projection"]
    out["last_normed"] = o
    return out


def frozen_openclip_embedder2(sd, tokens, heads, layer="last", return_pooled=False, legacy=True, prefix=""):
    """FrozenOpenCLIPEmbedder2.forward after tokenisation (modules.py:553-563)."""
    o = open_clip_text(sd, tokens, heads, prefix + "model.")
    if legacy:
        assert not return_pooled
        return _ln(sd, prefix + "model.ln_final", o[layer])          # modules.py:572-575
    return (o[layer], o["pooled"]) if return_pooled else o[layer]


def concat_timestep_embedder_nd(x, outdim):
    """modules.py:1027-1043."""
    if x.ndim == 1:
        x = x[:, None]
    b, dims = x.shape
    return timestep_embedding(x.reshape(-1), outdim).reshape(b, dims * outdim)


def supir_conditioner(sd, batch, heads_l, heads_g, clip_layer_idx=11, outdim=256, zero_keys=()):
    """GeneralConditionerWithControl.forward for the embedder list of options/SUPIR_v0.yaml:66-105 (modules.py:193-243):
    batch holds 'txt_tokens_l' / 'txt_tokens_g' (the two tokenisations of batch['txt']) and the three size tuples."""
    zl = frozen_clip_embedder(sd, batch["txt_tokens_l"], heads_l, "hidden", clip_layer_idx, prefix="embedders.0.")
    zg, pooled = frozen_openclip_embedder2(sd, batch["txt_tokens_g"], heads_g, "penultimate", True, False, prefix="embedders.1.")
    if "txt" in zero_keys:
        zl, zg, pooled = torch.zeros_like(zl), torch.zeros_like(zg), torch.zeros_like(pooled)
    vec = [pooled] + [concat_timestep_embedder_nd(batch[k], outdim) for k in
                      ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple")]
    return {"crossattn": torch.cat([zl, zg], 2), "vector": torch.cat(vec, 1), "control": batch["control"]}
