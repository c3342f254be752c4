"""Oracle: SDXL VAE encoder/decoder (untiled) and the tiled VAE of SUPIR/utils/tilevae.py (fp32, functional).

The network is first flattened into a list of steps (the reference's "task queue", tilevae.py:374-499); the untiled
forward runs the steps on one tensor with an ordinary GroupNorm, the tiled forward runs them on every padded tile and
replaces each GroupNorm by the reference's cross-tile statistic merge (tilevae.py:511-553, 599-648).
Test infrastructure only (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def _resblock_steps(sd, p):
    """resblock2task (tilevae.py:386-412) for sgm ResnetBlock (model.py:91-148)."""
    short = p + ".nin_shortcut" if (p + ".nin_shortcut.weight") in sd else None
    return [("store_res", short), ("norm", p + ".norm1"), ("silu",), ("conv", p + ".conv1"),
            ("norm", p + ".norm2"), ("silu",), ("conv", p + ".conv2"), ("add_res",)]


def _attn_steps(p):
    """attn2task (tilevae.py:349-372) for AttnBlock / MemoryEfficientAttnBlock (model.py:158-262)."""
    return [("store_res", None), ("norm", p + ".norm"), ("attn", p), ("add_res",)]


def build_steps(sd, prefix, is_decoder):
    """build_task_queue / build_sampling (tilevae.py:415-499); Encoder.forward / Decoder.forward (model.py:571-596, 710-743)."""
    p = prefix
    steps = [("conv", p + "conv_in")]
    mid = _resblock_steps(sd, p + "mid.block_1") + _attn_steps(p + "mid.attn_1") + _resblock_steps(sd, p + "mid.block_2")
    if is_decoder:
        steps += mid
        levels = _count(sd, p + "up")
        for lvl in reversed(range(levels)):
            for j in range(_count(sd, f"{p}up.{lvl}.block")):
                steps += _resblock_steps(sd, f"{p}up.{lvl}.block.{j}")
            if lvl != 0:
                steps.append(("upsample", f"{p}up.{lvl}.upsample.conv"))
    else:
        levels = _count(sd, p + "down")
        for lvl in range(levels):
            for j in range(_count(sd, f"{p}down.{lvl}.block")):
                steps += _resblock_steps(sd, f"{p}down.{lvl}.block.{j}")
            if lvl != levels - 1:
                steps.append(("downsample", f"{p}down.{lvl}.downsample.conv"))
        steps += mid
    steps += [("norm", p + "norm_out"), ("silu",), ("conv", p + "conv_out")]
    return steps


def _conv(sd, key, x, stride=1, padding=1):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def _attn(sd, p, h):
    """attn_forward (tilevae.py:292-315) == AttnBlock.attention + proj_out (model.py:177-200); scale = c^-0.5."""
    q, k, v = _conv(sd, p + ".q", h, padding=0), _conv(sd, p + ".k", h, padding=0), _conv(sd, p + ".v", h, padding=0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    out = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return _conv(sd, p + ".proj_out", out, padding=0)


def _apply_plain(sd, step, x):
    kind = step[0]
    if kind == "conv":
        return _conv(sd, step[1], x)
    if kind == "silu":
        return F.silu(x)
    if kind == "attn":
        return _attn(sd, step[1], x)
    if kind == "upsample":   # model.py:64-68
        return _conv(sd, step[1], F.interpolate(x, scale_factor=2.0, mode="nearest"))
    if kind == "downsample":  # model.py:81-85: pad right/bottom by one, stride 2, no conv padding
        return _conv(sd, step[1], F.pad(x, (0, 1, 0, 1), mode="constant", value=0), stride=2, padding=0)
    raise ValueError(kind)


def forward(sd, prefix, x, is_decoder):
    """Untiled Encoder.forward / Decoder.forward."""
    res = []
    for step in build_steps(sd, prefix, is_decoder):
        if step[0] == "store_res":
            res.append(x if step[1] is None else _conv(sd, step[1], x, padding=0))
        elif step[0] == "add_res":
            x = x + res.pop()
        elif step[0] == "norm":
            x = F.group_norm(x, 32, sd[step[1] + ".weight"], sd[step[1] + ".bias"], 1e-6)
        else:
            x = _apply_plain(sd, step, x)
    return x


# --------------------------------------------------------------------------------------------------------------------
# tiling bookkeeping (integer, must be bit-exact)
# --------------------------------------------------------------------------------------------------------------------
def get_best_tile_size(lowerbound, upperbound):
    """VAEHook.get_best_tile_size (tilevae.py:702-715)."""
    divider = 32
    while divider >= 2:
        remainer = lowerbound % divider
        if remainer == 0:
            return lowerbound
        candidate = lowerbound - remainer + divider
        if candidate <= upperbound:
            return candidate
        divider //= 2
    return lowerbound


def split_tiles(h, w, tile_size, is_decoder):
    """VAEHook.split_tiles (tilevae.py:717-774); pad = 11 (decoder) / 32 (encoder) (tilevae.py:686). bbox = [x1,x2,y1,y2]."""
    pad = 11 if is_decoder else 32
    nh = max(math.ceil((h - 2 * pad) / tile_size), 1)
    nw = max(math.ceil((w - 2 * pad) / tile_size), 1)
    real_h = get_best_tile_size(math.ceil((h - 2 * pad) / nh), tile_size)
    real_w = get_best_tile_size(math.ceil((w - 2 * pad) / nw), tile_size)
    in_bboxes, out_bboxes = [], []
    for i in range(nh):
        for j in range(nw):
            ib = [pad + j * real_w, min(pad + (j + 1) * real_w, w), pad + i * real_h, min(pad + (i + 1) * real_h, h)]
            ob = [ib[0] if ib[0] > pad else 0, ib[1] if ib[1] < w - pad else w,
                  ib[2] if ib[2] > pad else 0, ib[3] if ib[3] < h - pad else h]
            out_bboxes.append([v * 8 if is_decoder else v // 8 for v in ob])
            in_bboxes.append([max(0, ib[0] - pad), min(w, ib[1] + pad), max(0, ib[2] - pad), min(h, ib[3] + pad)])
    return in_bboxes, out_bboxes


def crop_margins(tile_h, tile_w, input_bbox, target_bbox, is_decoder):
    """crop_valid_region (tilevae.py:556-567) as index ranges: rows [y0, y1), cols [x0, x1) of the tile output."""
    padded = [i * 8 if is_decoder else i // 8 for i in input_bbox]
    margin = [target_bbox[i] - padded[i] for i in range(4)]
    return margin[2], tile_h + margin[3], margin[0], tile_w + margin[1]


def needs_tiling(h, w, tile_size, is_decoder):
    """VAEHook.__call__ (tilevae.py:688-700)."""
    pad = 11 if is_decoder else 32
    return max(h, w) > pad * 2 + tile_size


# --------------------------------------------------------------------------------------------------------------------
# tiled forward
# --------------------------------------------------------------------------------------------------------------------
def _var_mean(t):
    """get_var_mean (tilevae.py:511-521): biased variance and mean per (image, group)."""
    b, c = t.shape[0], t.shape[1]
    r = t.contiguous().view(1, b * 32, c // 32, *t.shape[2:])
    return torch.var_mean(r, dim=[0, 2, 3, 4], unbiased=False)


def _custom_group_norm(t, mean, var, weight, bias, eps=1e-6):
    """custom_group_norm (tilevae.py:524-553)."""
    b, c = t.shape[0], t.shape[1]
    r = t.contiguous().view(1, b * 32, c // 32, *t.shape[2:])
    out = F.batch_norm(r, mean, var, weight=None, bias=None, training=False, momentum=0, eps=eps).view(b, c, *t.shape[2:])
    return out * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def fast_mode_thumbnail(z, tile_size):
    """vae_tile_forward, fast mode (tilevae.py:855-873): nearest-exact downsample of the whole input to about one tile, its
    per-channel mean / std restored to the full input's, clamped to the input's range."""
    scale_factor = tile_size / max(z.shape[2], z.shape[3])
    d = F.interpolate(z, scale_factor=scale_factor, mode="nearest-exact")
    std_old, mean_old = torch.std_mean(z, dim=[0, 2, 3], keepdim=True)
    std_new, mean_new = torch.std_mean(d, dim=[0, 2, 3], keepdim=True)
    d = (d - mean_new) / std_new * std_old + mean_old
    return torch.clamp_(d, min=z.min(), max=z.max())


def estimate_group_norm(sd, steps, x, color_fix=False):
    """estimate_group_norm (tilevae.py:776-817): run the layers on the thumbnail as ONE tile and keep the (var, mean) each
    GroupNorm sees there (GroupNormParam.from_tile); with color_fix only the layers before the first downsample are estimated.
    Returns one (var, mean) or None per 'norm' step."""
    res, fixed = [], []
    n_norm = sum(1 for s_ in steps if s_[0] == "norm")
    for step in steps:
        if step[0] == "norm":
            var, mean = _var_mean(x)
            fixed.append((var, mean))
            if len(fixed) == n_norm:
                break
            x = _custom_group_norm(x, mean, var, sd[step[1] + ".weight"], sd[step[1] + ".bias"])
        elif step[0] == "store_res":
            res.append(x if step[1] is None else _conv(sd, step[1], x, padding=0))
        elif step[0] == "add_res":
            x = x + res.pop()
        elif color_fix and step[0] == "downsample":
            break
        else:
            x = _apply_plain(sd, step, x)
    return fixed + [None] * (n_norm - len(fixed))


def tiled_forward(sd, prefix, z, tile_size, is_decoder, fast=False, color_fix=False):
    """VAEHook.vae_tile_forward (tilevae.py:819-970); fast=True is its fast mode (statistics from estimate_group_norm, no
    cross-tile merge for the estimated layers)."""
    n, _, height, width = z.shape
    if not needs_tiling(height, width, tile_size, is_decoder):
        return forward(sd, prefix, z, is_decoder)
    in_bboxes, out_bboxes = split_tiles(height, width, tile_size, is_decoder)
    tiles = [z[:, :, b[2]:b[3], b[0]:b[1]] for b in in_bboxes]
    res = [[] for _ in tiles]
    steps = build_steps(sd, prefix, is_decoder)
    fixed = estimate_group_norm(sd, steps, fast_mode_thumbnail(z, tile_size), color_fix and not is_decoder) if fast else None
    norm_idx = -1
    for step in steps:
        if step[0] == "norm":
            norm_idx += 1
            if fixed is not None and fixed[norm_idx] is not None:
                var, mean = fixed[norm_idx]
                tiles = [_custom_group_norm(t, mean, var, sd[step[1] + ".weight"], sd[step[1] + ".bias"]) for t in tiles]
                continue
            stats = [_var_mean(t) for t in tiles]
            pixels = torch.tensor([t.shape[2] * t.shape[3] for t in tiles], dtype=torch.float32)
            pixels = pixels / pixels.max()
            wts = (pixels / pixels.sum()).unsqueeze(1)                         # GroupNormParam.summary (tilevae.py:629-648)
            var = torch.sum(torch.vstack([s[0] for s in stats]) * wts, dim=0)
            mean = torch.sum(torch.vstack([s[1] for s in stats]) * wts, dim=0)
            tiles = [_custom_group_norm(t, mean, var, sd[step[1] + ".weight"], sd[step[1] + ".bias"]) for t in tiles]
        elif step[0] == "store_res":
            for i, t in enumerate(tiles):
                res[i].append(t if step[1] is None else _conv(sd, step[1], t, padding=0))
        elif step[0] == "add_res":
            tiles = [t + res[i].pop() for i, t in enumerate(tiles)]
        else:
            tiles = [_apply_plain(sd, step, t) for t in tiles]
    out_c = tiles[0].shape[1]
    result = torch.zeros((n, out_c, height * 8 if is_decoder else height // 8, width * 8 if is_decoder else width // 8))
    for t, ib, ob in zip(tiles, in_bboxes, out_bboxes):
        y0, y1, x0, x1 = crop_margins(t.shape[2], t.shape[3], ib, ob, is_decoder)
        result[:, :, ob[2]:ob[3], ob[0]:ob[1]] = t[:, :, y0:y1, x0:x1]
    return result


# --------------------------------------------------------------------------------------------------------------------
# AutoencoderKL wrappers (sgm/models/autoencoder.py:304-321; SUPIR/models/SUPIR_model.py:41-69)
# --------------------------------------------------------------------------------------------------------------------
def encode_moments(sd, x, encoder_prefix="encoder.", tile_size=None):
    h = tiled_forward(sd, encoder_prefix, x, tile_size, False) if tile_size else forward(sd, encoder_prefix, x, False)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def gaussian_latent(moments, eps=None, scale_factor=0.13025):
    """DiagonalGaussianDistribution (distributions.py:24-41): eps=None -> mode(); then scale (SUPIR_model.py:45,61)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    z = mean if eps is None else mean + torch.exp(0.5 * logvar) * eps
    return scale_factor * z


def decode(sd, z, scale_factor=0.13025, decoder_prefix="decoder.", tile_size=None):
    z = 1.0 / scale_factor * z
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    out = tiled_forward(sd, decoder_prefix, z, tile_size, True) if tile_size else forward(sd, decoder_prefix, z, True)
    return out.float()
