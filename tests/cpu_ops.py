"""Plain-torch CPU stand-ins for the few supir_b200.ops kernels the SAMPLERS call (test infrastructure: lets the host-side
sampler logic — step constants, unit sharding, exchanges, run objects — be exercised without a GPU, including under gloo).
Never imported by the product."""
import torch


def install(monkeypatch_or_none=None):
    from supir_b200 import ops

    def tile_gather(src, windows, tile, out):
        for j, (hi, he, wi, we) in enumerate(windows.tolist()):
            if hi >= 0:
                out[j] = src[:, :, hi:he, wi:we]
        return out

    def tile_blend(tiles, windows, tile, weights, out):
        acc = torch.zeros(out.shape, dtype=torch.float32)
        cnt = torch.zeros(out.shape, dtype=torch.float32)
        w = weights.to(torch.float64)
        for j, (hi, he, wi, we) in enumerate(windows.tolist()):
            if hi < 0:
                continue
            acc[:, :, hi:he, wi:we] += (tiles[j].to(torch.float64) * w).to(torch.float32)
            cnt[:, :, hi:he, wi:we] += w.to(torch.float32)
        out.copy_(acc / cnt)
        return out

    def edm_pre(x, eps, noise_mul, c_in, x_hat, net_in):
        v = x if eps is None else x + eps * noise_mul
        x_hat.copy_(v)
        net_in.view(2, -1).copy_((v * c_in).reshape(1, -1).expand(2, -1))

    def edm_post(x_hat, net_out, x_center, c_out, cfg_scale, restore_mul, sigma_hat, dt, x_next, denoised=None):
        n = x_hat.numel()
        net = net_out.reshape(2, n)
        xh = x_hat.reshape(-1)
        du, dc = net[0] * c_out + xh, net[1] * c_out + xh
        den = du + cfg_scale * (dc - du)
        if x_center is not None:
            den = den - (den - x_center.reshape(-1)) * restore_mul
        x_next.copy_((xh + (xh - den) / sigma_hat * dt).view(x_next.shape))
        if denoised is not None:
            denoised.copy_(den.view(denoised.shape))

    def axpby_f32(a, alpha, b, beta, out):
        out.copy_(a * alpha if b is None else a * alpha + b * beta)
        return out

    def cfg_combine(x, scale, out):
        u, c = x.chunk(2)
        out.copy_(u + scale.view(-1, *([1] * (x.dim() - 1))) * (c - u))
        return out

    fns = dict(tile_gather=tile_gather, tile_blend=tile_blend, edm_pre=edm_pre, edm_post=edm_post, axpby_f32=axpby_f32, cfg_combine=cfg_combine)
    for k, f in fns.items():
        if monkeypatch_or_none is not None:
            monkeypatch_or_none.setattr(ops, k, f)
        else:
            setattr(ops, k, f)
