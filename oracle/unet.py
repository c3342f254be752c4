"""Oracle: GLVControl + LightGLVUNet forward (fp32, functional, driven by the reference's state_dict keys).

The block layout is recovered from the key names themselves, so the same code serves the full SDXL-base configuration
and the tiny configurations used for fixtures. Test infrastructure only (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------------------------
def timestep_embedding(timesteps, dim, max_period=10000):
    """sgm/modules/diffusionmodules/util.py:206-230 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _has(sd, key):
    return key in sd


def embed(sd, p, timesteps, y, model_channels):
    """time_embed MLP + label_emb MLP (openaimodel.py:993-998; SUPIR_v0.py:515-522, 618-623)."""
    t_emb = timestep_embedding(timesteps, model_channels)
    emb = _lin(sd, p + "time_embed.2", F.silu(_lin(sd, p + "time_embed.0", t_emb)))
    emb = emb + _lin(sd, p + "label_emb.0.2", F.silu(_lin(sd, p + "label_emb.0.0", y)))
    return emb


def resblock(sd, p, x, emb):
    """ResBlock._forward, no up/down, no scale-shift (openaimodel.py:330-356)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if _has(sd, p + ".skip_connection.weight"):
        x = _conv(sd, p + ".skip_connection", x, padding=0)
    return x + h


def attention(sd, p, x, context, heads):
    """CrossAttention.forward (attention.py:222-285): softmax(q k^T d^-0.5) v, to_out."""
    ctx = x if context is None else context
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    b, n, c = q.shape
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    out = F.scaled_dot_product_attention(q, k, v)
    out = out.transpose(1, 2).reshape(b, n, c)
    return _lin(sd, p + ".to_out.0", out)


def transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward (attention.py:465-486), GEGLU feed-forward (attention.py:84-110)."""
    x = attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads) + x
    x = attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads) + x
    h = _lin(sd, p + ".ff.net.0.proj", _ln(sd, p + ".norm3", x))
    a, gate = h.chunk(2, dim=-1)
    x = _lin(sd, p + ".ff.net.2", a * F.gelu(gate)) + x
    return x


def spatial_transformer(sd, p, x, context, head_dim):
    """SpatialTransformer.forward with use_linear=True (attention.py:614-635); Normalize eps 1e-6 (attention.py:122-125)."""
    b, c, h, w = x.shape
    x_in = x
    t = _gn(sd, p + ".norm", x, 1e-6).permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = _lin(sd, p + ".proj_in", t)
    heads = t.shape[-1] // head_dim
    d = 0
    while _has(sd, f"{p}.transformer_blocks.{d}.norm1.weight"):
        t = transformer_block(sd, f"{p}.transformer_blocks.{d}", t, context, heads)
        d += 1
    t = _lin(sd, p + ".proj_out", t)
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2) + x_in


def _run_block(sd, p, h, emb, context, head_dim, stop_before_upsample=False):
    """TimestepEmbedSequential.forward (openaimodel.py:87-105) over whatever sub-layers the keys reveal."""
    i = 0
    while True:
        q = f"{p}.{i}"
        if _has(sd, q + ".in_layers.0.weight"):
            h = resblock(sd, q, h, emb)
        elif _has(sd, q + ".proj_in.weight"):
            h = spatial_transformer(sd, q, h, context, head_dim)
        elif _has(sd, q + ".op.weight"):  # Downsample (openaimodel.py:170-210): 3x3 stride 2 pad 1
            h = _conv(sd, q + ".op", h, stride=2, padding=1)
        elif _has(sd, q + ".conv.weight"):  # Upsample (openaimodel.py:108-151): nearest 2x then 3x3
            if stop_before_upsample:
                return h, q
            h = _conv(sd, q + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"))
        elif _has(sd, q + ".weight"):  # bare conv (input_blocks.0)
            h = _conv(sd, q, h)
        else:
            break
        i += 1
    return (h, None) if stop_before_upsample else h


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


# --------------------------------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------------------------------
def glv_control_forward(sd, x, timesteps, xt, context, y, model_channels, head_dim, prefix=""):
    """GLVControl.forward (SUPIR/modules/SUPIR_v0.py:499-540). Returns the list of feature maps `hs`."""
    p = prefix
    emb = embed(sd, p, timesteps, y, model_channels)
    guided_hint = _conv(sd, p + "input_hint_block.0", x)
    hs = []
    h = xt
    for i in range(_count(sd, p + "input_blocks")):
        h = _run_block(sd, f"{p}input_blocks.{i}", h, emb, context, head_dim)
        if guided_hint is not None:
            h = h + guided_hint
            guided_hint = None
        hs.append(h)
    h = _run_block(sd, p + "middle_block", h, emb, context, head_dim)
    hs.append(h)
    return hs


def zero_sft(sd, p, c, h, h_ori=None, control_scale=1.0):
    """ZeroSFT.forward (SUPIR_v0.py:91-113); pre_concat is True whenever concat_channels != 0."""
    pre_concat = sd[p + ".param_free_norm.weight"].shape[0] != sd[p + ".zero_conv.weight"].shape[0]
    if h_ori is not None and pre_concat:
        h_raw = torch.cat([h_ori, h], dim=1)
    else:
        h_raw = h
    h = h + _conv(sd, p + ".zero_conv", c, padding=0)
    if h_ori is not None and pre_concat:
        h = torch.cat([h_ori, h], dim=1)
    actv = F.silu(_conv(sd, p + ".mlp_shared.0", c))
    gamma = _conv(sd, p + ".zero_mul", actv)
    beta = _conv(sd, p + ".zero_add", actv)
    h = _gn(sd, p + ".param_free_norm", h, 1e-5) * (gamma + 1) + beta
    if h_ori is not None and not pre_concat:
        h = torch.cat([h_ori, h], dim=1)
    return h * control_scale + h_raw * (1 - control_scale)


def zero_cross_attn(sd, p, context, x, control_scale=1.0):
    """ZeroCrossAttn.forward (SUPIR_v0.py:138-152): heads = query_dim // 64."""
    x_in = x
    b, c, h, w = x.shape
    xn = _gn(sd, p + ".norm1", x, 1e-5).permute(0, 2, 3, 1).reshape(b, h * w, c)
    cn = _gn(sd, p + ".norm2", context, 1e-5)
    cn = cn.permute(0, 2, 3, 1).reshape(b, -1, context.shape[1])
    heads = sd[p + ".attn.to_q.weight"].shape[0] // 64
    a = attention(sd, p + ".attn", xn, cn, max(heads, 1))
    a = a.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x_in + a * control_scale


def light_glv_unet_forward(sd, x, timesteps, context, y, control, control_scale, model_channels, head_dim, prefix=""):
    """LightGLVUNet.forward (SUPIR_v0.py:600-666)."""
    p = prefix
    emb = embed(sd, p, timesteps, y, model_channels)
    hs = []
    h = x
    for i in range(_count(sd, p + "input_blocks")):
        h = _run_block(sd, f"{p}input_blocks.{i}", h, emb, context, head_dim)
        hs.append(h)

    def project(idx, *args, **kw):
        q = f"{p}project_modules.{idx}"
        if _has(sd, q + ".attn.to_q.weight"):
            return zero_cross_attn(sd, q, *args, **kw)
        return zero_sft(sd, q, *args, **kw)

    adapter_idx = _count(sd, p + "project_modules") - 1
    control_idx = len(control) - 1
    h = _run_block(sd, p + "middle_block", h, emb, context, head_dim)
    h = project(adapter_idx, control[control_idx], h, control_scale=control_scale)
    adapter_idx -= 1
    control_idx -= 1
    for i in range(_count(sd, p + "output_blocks")):
        _h = hs.pop()
        h = project(adapter_idx, control[control_idx], _h, h, control_scale=control_scale)
        adapter_idx -= 1
        ob = f"{p}output_blocks.{i}"
        if _count(sd, ob) == 3:
            h, up = _run_block(sd, ob, h, emb, context, head_dim, stop_before_upsample=True)
            h = project(adapter_idx, control[control_idx], h, control_scale=control_scale)
            adapter_idx -= 1
            h = _conv(sd, up + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"))
        else:
            h = _run_block(sd, ob, h, emb, context, head_dim)
        control_idx -= 1
    return _conv(sd, p + "out.2", F.silu(_gn(sd, p + "out.0", h, 1e-5)))


def control_wrapper_forward(sd, x, t, c, control_scale, model_channels=320, head_dim=64,
                            unet_prefix="diffusion_model.", control_prefix="control_model."):
    """ControlWrapper.forward (sgm/modules/diffusionmodules/wrappers.py:84-102) in fp32."""
    control = glv_control_forward(sd, c["control"], t, x, c["crossattn"], c["vector"], model_channels, head_dim,
                                  prefix=control_prefix)
    out = light_glv_unet_forward(sd, x, t, c["crossattn"], c["vector"], control, control_scale, model_channels,
                                 head_dim, prefix=unet_prefix)
    return out.float()
