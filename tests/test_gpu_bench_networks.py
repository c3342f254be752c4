"""GPU parity of the NETWORKS at the configurations bench.py runs, against the CPU oracle (fp32 restatement of the reference):

  * the real SUPIR-v0 / SDXL-base layout (transformer depth [1, 2, 10], 3.87 B parameters) on one 128x128 latent window,
    CFG pair — the unit of work of BASELINE configs[2] (one of its 49 windows), where the tile dispatcher picks the wide
    2-CTA tiles and attention runs 8 / 32 key blocks;
  * BASELINE configs[0] ("cfg1"): 64x64 latent, ONE EDM step through RestoreEDMSampler (sigma 14.6146 -> 0) with churn noise,
    LinearCFG and control scale, vs the oracle's sampler on the same weights / inputs / noise;
  * the real SDXL VAE (ch 128, 512-wide single-head mid attention) on ONE padded encoder tile of the bench's tiling
    (1024 px + 2 x 32 -> 1088 px, 136^2 = 18 496 attention tokens) and ONE padded decoder tile (128 + 2 x 11 -> 150 latent,
    22 500 tokens);
  * error growth over the 50 EDM steps of the benchmark (depth-1 full-width network, 16x16 latent): relative error vs the
    oracle trajectory recorded at steps 1 / 10 / 25 / 50.

Tolerance: bf16 storage with fp32 accumulation against fp32 — relative Frobenius error <= 3e-2 per network call (measured
~1e-2), <= 3e-2 for a VAE tile, <= 8e-2 after 50 sampler steps (the trajectory amplifies per-call error; see the printed
growth curve). Reference call sites: SUPIR/modules/SUPIR_v0.py:499-540,600-666; sgm/modules/diffusionmodules/sampling.py:
548-597; sgm/modules/diffusionmodules/model.py:571-596,710-743."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from weights import make_state_dict, randn, shapes_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}


def rel_fro(a, b):
    return float((a - b).norm() / b.norm())


def _threads():
    torch.set_num_threads(min(os.cpu_count() or 1, 64))


@pytest.fixture(scope="module")
def full_depth():
    """(state_dict, ControlWrapper on cuda) of the full SUPIR-v0 layout with random weights (non-trivial biases / norms)."""
    sys.path.insert(0, ROOT)
    import bench
    from supir_b200 import nets, wrappers
    sd = bench.oracle_state_dict()
    g = torch.Generator().manual_seed(3)
    for k, v in sd.items():
        if not (k.endswith("weight") and v.dim() >= 2):
            v.add_(0.05 * torch.randn(v.shape, generator=g))
    with torch.device("meta"):
        unet = nets.LightGLVUNet(mode="XL-base", project_type="ZeroSFT", project_channel_scale=2, **bench.UNET_CFG)
        ctrl = nets.GLVControl(input_upscale=1, **bench.UNET_CFG)
    w = wrappers.ControlWrapper(unet, dtype=torch.bfloat16)
    w.load_control_model(ctrl)
    w.to_empty(device="cuda")
    w.load_state_dict(sd, strict=True)
    _threads()
    yield sd, w
    w.invalidate()
    torch.cuda.empty_cache()


def _cond(side, seed, n=2):
    return {"control": randn((n, 4, side, side), seed), "crossattn": randn((n, 77, 2048), seed + 1), "vector": randn((n, 2816), seed + 2)}


@pytest.mark.slow
@pytest.mark.parametrize("side,tol", [(32, 3e-2), (128, 3e-2)])
def test_full_depth_control_unet_vs_oracle(full_depth, side, tol):
    """One denoiser-network call (GLVControl + LightGLVUNet, CFG pair) at a 32^2 latent and at the bench's 128^2 window."""
    from oracle import unet as ounet
    sd, w = full_depth
    x, cond = randn((2, 4, side, side), 301 + side), _cond(side, 310 + side)
    t = torch.tensor([700, 700])
    out = w(x.cuda(), t.cuda(), {k: v.cuda() for k, v in cond.items()}, control_scale=1.0).cpu()
    ref = ounet.control_wrapper_forward(sd, x, t, cond, 1.0)
    e, m = rel_fro(out, ref), float((out - ref).abs().max() / ref.abs().max())
    print(f"full-depth control+UNet at {side}x{side} latent: rel_fro={e:.4g} max_rel={m:.4g}")
    assert e <= tol and m <= 3 * tol
    w.invalidate()            # drop the plan's activation pool before the next (larger) shape


@pytest.mark.slow
def test_cfg1_one_edm_step_64_latent_vs_oracle(full_depth):
    """BASELINE configs[0]: 64x64 latent (512 px), one EDM step, s_churn 5, s_noise 1.01, LinearCFG(1.0 -> 4.0), restore_cfg -1."""
    from oracle import sampler as osamp, unet as ounet
    from supir_b200 import denoiser as dn, sampling
    sd, w = full_depth
    den = dn.DiscreteDenoiserWithControl(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config=DISC).cuda()
    smp = sampling.RestoreEDMSampler(
        num_steps=1, restore_cfg=-1.0, s_churn=5, s_noise=1.01, discretization_config=DISC,
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 1.0, "scale_min": 4.0}})
    x = randn((1, 4, 64, 64), 401)
    c, uc = _cond(64, 410, 1), _cond(64, 420, 1)
    uc["control"] = c["control"]
    xc = randn((1, 4, 64, 64), 430)
    eps = randn((1, 4, 64, 64), 440)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    sig = smp.host_sigmas()
    assert len(sig) == 2 and abs(float(sig[0]) - 14.6146) < 1e-3 and float(sig[1]) == 0.0
    x0, sigmas = smp.prepare_sampling_loop(x.cuda())
    k = smp.step_constants(sigmas, 0, 1.0, False, 0.0)
    out = smp._step(sampling.FusedDenoiser(den, w), x0, eps.cuda(), cu(c), cu(uc), xc.cuda(), k).cpu()
    osm = osamp.RestoreEDMSampler(num_steps=1, restore_cfg=-1.0, s_churn=5, s_noise=1.01, scale=1.0, scale_min=4.0)
    net = lambda a, t, cc, cs: ounet.control_wrapper_forward(sd, a, t, cc, cs)  # noqa: E731
    xr, s_in, osig = osm.prepare(x)
    ref = osm.sampler_step(net, s_in * osig[0], s_in * osig[1], xr, c, uc, osm.gamma(osig, 0), xc, eps_noise=eps, control_scale=1.0)
    e = rel_fro(out, ref)
    print(f"cfg1 (64x64 latent, 1 EDM step): rel_fro={e:.4g}")
    assert e <= 3e-2
    w.invalidate()


def test_real_sdxl_vae_bench_tiles_vs_oracle():
    """The bench's VAE configuration on the padded tile sizes its tiling produces (bench.py: encoder tile 1024 px, decoder
    tile 128 latent): conv / GroupNorm at 128..512 channels and the 512-wide mid attention over 18 496 / 22 500 tokens."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import vae as ovae
    from supir_b200 import vae
    with torch.device("meta"):
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=bench.VAE_CFG, lossconfig={"target": "torch.nn.Identity"})
    sd = make_state_dict(shapes_of(ae), seed=91)
    ae.to_empty(device="cuda")
    ae.load_state_dict(sd, strict=True)
    _threads()
    in_b, _ = vae.split_tiles(4096, 4096, 1024, False)
    assert (in_b[5][1] - in_b[5][0], in_b[5][3] - in_b[5][2]) == (1088, 1088)
    img = randn((1, 3, 1088, 1088), 92) * 0.5
    got = ae.encoder.original_forward(img.cuda()).cpu()
    ref = ovae.forward(sd, "encoder.", img, False)
    e1 = rel_fro(got, ref)
    in_d, _ = vae.split_tiles(512, 512, 128, True)
    assert (in_d[5][1] - in_d[5][0], in_d[5][3] - in_d[5][2]) == (150, 150)
    z = randn((1, 4, 150, 150), 93)
    got = ae.decoder.original_forward(z.cuda()).cpu()
    ref = ovae.forward(sd, "decoder.", z, True)
    e2 = rel_fro(got, ref)
    print(f"real SDXL VAE tiles: encoder 1088 px rel_fro={e1:.4g}, decoder 150 latent rel_fro={e2:.4g}")
    assert e1 <= 3e-2 and e2 <= 3e-2


def test_error_growth_over_50_edm_steps():
    """GPU trajectory vs oracle trajectory with identical noise: the per-call bf16 error must not blow up over the 50 steps
    of the benchmarked schedule. Records the curve (printed; the driver's log keeps it)."""
    from oracle import sampler as osamp, unet as ounet
    from supir_b200 import denoiser as dn, nets, sampling, wrappers
    g = np.load(os.path.join(G, "unet_fullwidth_depth1.npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = make_state_dict(json.loads(str(g["shapes"])), seed=31)
    with torch.device("meta"):
        unet = nets.LightGLVUNet(mode="XL-base", project_type="ZeroSFT", project_channel_scale=2, **cfg)
        ctrl = nets.GLVControl(input_upscale=1, **cfg)
    w = wrappers.ControlWrapper(unet, dtype=torch.bfloat16)
    w.load_control_model(ctrl)
    w.to_empty(device="cuda")
    w.load_state_dict(sd, strict=True)
    _threads()
    den = dn.DiscreteDenoiserWithControl(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config=DISC).cuda()
    steps = 50
    smp = sampling.RestoreEDMSampler(
        num_steps=steps, restore_cfg=4.0, s_churn=5, s_noise=1.01, discretization_config=DISC,
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 1.0, "scale_min": 4.0}})
    osm = osamp.RestoreEDMSampler(num_steps=steps, restore_cfg=4.0, s_churn=5, s_noise=1.01, scale=1.0, scale_min=4.0)
    side = 16
    x = randn((1, 4, side, side), 501)
    c, uc = _cond(side, 510, 1), _cond(side, 520, 1)
    c["control"] = c["control"] * 0.5
    uc["control"] = c["control"]
    xc = randn((1, 4, side, side), 530) * 0.5
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    cg, ucg, xcg = cu(c), cu(uc), xc.cuda()
    fd = sampling.FusedDenoiser(den, w)
    net = lambda a, t, cc, cs: ounet.control_wrapper_forward(sd, a, t, cc, cs)  # noqa: E731
    xg, sigmas = smp.prepare_sampling_loop(x.cuda())
    xr, s_in, osig = osm.prepare(x)
    curve = {}
    for i in range(steps):
        eps = randn((1, 4, side, side), 600 + i)
        k = smp.step_constants(sigmas, i, 0.9, True, 0.0)
        xg = smp._step(fd, xg, eps.cuda() if k["gamma"] > 0 else None, cg, ucg, xcg, k)
        xr = osm.sampler_step(net, s_in * osig[i], s_in * osig[i + 1], xr, c, uc, osm.gamma(osig, i), xc, eps_noise=eps,
                              control_scale=0.9, use_linear_control_scale=True, control_scale_start=0.0)
        if i + 1 in (1, 10, 25, 50):
            curve[i + 1] = rel_fro(xg.cpu(), xr)
    print("error growth over 50 EDM steps (rel. Frobenius, GPU bf16 vs oracle fp32): " + json.dumps(curve))
    assert all(np.isfinite(v) for v in curve.values())
    assert curve[1] <= 2e-2 and curve[50] <= 8e-2
