#!/usr/bin/env python
"""bench.py — SUPIR EDM sampling hot path on B200: megapixels/sec and ms/EDM-step, 1024^2 -> 4096^2 @ 50 steps.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU): one 1024x1024 image
restored at upscale 4 -> 4096x4096 output (16.78 MP), latent 512x512, `TiledRestoreEDMSampler(tile 128, stride 64)` = 49
windows per step (the gradio_demo_tiled.py reading of the config, the one that shards over GPUs), cond+uncond CFG pair,
s_churn 5, s_noise 1.01, linear CFG 1.0 -> 4.0, tiled VAE (encoder tile 1024 px, decoder tile 128 latent). Synthetic
input, random-init weights of the exact SUPIR-v0 / SDXL-base / SDXL-VAE architecture.

A "step" is ONE EDM sampler step: all 49 windows x (control net + UNet) on the CFG pair, the step arithmetic, the window
blend (and, for N > 1, the all-gather of window outputs). `ms_per_step` is its device time. `value` (megapixels/sec) is
16.777 MP / (50 * ms_per_step + the VAE passes (2 encodes + 2 decodes) timed in the same run). `e2e` repeats both with
host buffers in the loop (pinned host -> device copy of the step's latent and device -> host read-back of the result every
step; image upload + decoded image download around the VAE). N > 1 shards the windows of the SAME image over ranks
(strong scaling). `--impl reference` times the CPU oracle (the reference algorithm) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_CFG = dict(adm_in_channels=2816, num_classes="sequential", use_checkpoint=True, in_channels=4, out_channels=4,
                model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4],
                num_head_channels=64, use_spatial_transformer=True, use_linear_in_transformer=True,
                transformer_depth=[1, 2, 10], context_dim=2048, spatial_transformer_attn_type="softmax-xformers", legacy=False)
VAE_CFG = dict(attn_type="vanilla-xformers", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
               ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
OUT_PX, LATENT, TILE, STRIDE, EDM_STEPS = 4096, 512, 128, 64, 50
MEGAPIXELS = OUT_PX * OUT_PX / 1e6
# algorithmic FLOP of one denoiser call (control + UNet, CFG pair) on one 128x128 window: SURVEY.md §8(d), BASELINE.md §3
FLOP_WINDOW = 20.292e12
FLOP_BY_LATENT = {64: 4.766e12, 128: 20.292e12, 256: 107.887e12}


def flop_denoiser(side):
    if side in FLOP_BY_LATENT:
        return FLOP_BY_LATENT[side]
    hw = side * side
    return 1.1e9 * hw + 0.1416e12 * (hw / 4096.0) ** 2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["bf16_tflops"], p["hbm_gbs"], "measured"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch_mean"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.lines, self.i0 = index, None, [], 0

    def mark(self):
        """Start of the timed region: samples read from here on are the ones reported. The process is started before the
        warm-up steps because nvidia-smi needs up to a second to emit its first line on a multi-GPU box."""
        self.i0 = len(self.lines)

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        window = "timed region"
        lines = self.lines[self.i0:]
        if not any(len(ln.split(",")) >= 7 for ln in lines):
            lines, window = self.lines[-8:], "warm-up + timed steps (same load; no sample fell inside the short timed region)"
        sm, mx, reasons = [], None, set()
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def oracle_state_dict():
    """Random fp32 weights with the full SUPIR-v0 shapes, keyed like the reference checkpoint (model.* prefix dropped)."""
    from supir_b200 import nets
    with torch.device("meta"):
        unet = nets.LightGLVUNet(mode="XL-base", project_type="ZeroSFT", project_channel_scale=2, **UNET_CFG)
        ctrl = nets.GLVControl(input_upscale=1, **UNET_CFG)
    sd = {}
    g = torch.Generator().manual_seed(0)
    for prefix, m in (("diffusion_model.", unet), ("control_model.", ctrl)):
        for k, v in m.state_dict().items():
            t = torch.empty(v.shape, dtype=torch.float32)
            if k.endswith("weight") and v.dim() >= 2:
                fan_in = int(np.prod(v.shape[1:]))
                t.normal_(0, 1.0 / fan_in ** 0.5, generator=g)
            elif k.endswith("weight"):
                t.fill_(1.0)
            else:
                t.zero_()
            sd[prefix + k] = t
    return sd


def time_oracle_window(sd, side, repeats):
    """Seconds per denoiser call (control + UNet on the CFG pair, B=2) on one side x side latent window, fp32 CPU."""
    from oracle import unet as ounet
    x = torch.randn(2, 4, side, side)
    cond = {"control": torch.randn(2, 4, side, side), "crossattn": torch.randn(2, 77, 2048), "vector": torch.randn(2, 2816)}
    t = torch.tensor([500, 500])
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            ounet.control_wrapper_forward(sd, x, t, cond, 1.0)
            times.append(time.perf_counter() - t0)
    return times


def mp_per_s_from_window_time(sec_128):
    """Whole-job throughput from the time of one 128-latent window step: 49 windows x 50 steps (the VAE, ~1.4 % of the
    job's work, is left out: this flatters the CPU)."""
    return MEGAPIXELS / (sec_128 * 49 * EDM_STEPS), sec_128 * 49 * 1e3


def probe_oracle(budget_s, sd, repeats=1):
    """Time the oracle's denoiser call on growing windows (latent side 16, 32, 64, 128) while a call fits the budget.
    Returns ({side: seconds}, estimated seconds for the 128-latent window). When 128 itself was not reached the estimate
    is a two-point fit t = a + b * FLOP(side) through the two largest measured sizes (a = the size-independent cost of
    streaming 15.5 GB of fp32 weights), which is fairer to the CPU than scaling a tiny window by FLOPs alone."""
    time_oracle_window(sd, 16, 1)        # warm-up: first touch of the weights, oneDNN primitive creation
    meas, side = {}, 16
    while True:
        meas[side] = float(np.mean(time_oracle_window(sd, side, repeats)))
        if side >= 128 or meas[side] * 4.5 > budget_s:
            break
        side *= 2
    sides = sorted(meas)
    if 128 in meas:
        return meas, meas[128]
    if len(sides) == 1:
        return meas, meas[sides[0]] * FLOP_WINDOW / flop_denoiser(sides[0])
    s0, s1 = sides[-2], sides[-1]
    b = (meas[s1] - meas[s0]) / (flop_denoiser(s1) - flop_denoiser(s0))
    a = max(meas[s1] - b * flop_denoiser(s1), 0.0)
    return meas, a + max(b, 0.0) * FLOP_WINDOW


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, int(os.environ.get("SUPIR_BENCH_CPU_THREADS", "64")))
    torch.set_num_threads(cores)
    sd = oracle_state_dict()
    total_budget = float(os.environ.get("SUPIR_BENCH_REF_BUDGET_S", "150"))
    reps = max(args.steps, 1)
    meas, sec_128 = probe_oracle(total_budget / max(args.steps + args.warmup, 1), sd, repeats=reps)
    side = max(meas)
    sec = meas[side]
    mps, ms_step = mp_per_s_from_window_time(sec_128)
    sample = (f"{reps} timed oracle calls per size (fp32 torch restatement of the reference, {cores} threads) of control+UNet on ONE "
              f"window (CFG pair) at latent sides {sorted(meas)}: {[round(meas[k], 2) for k in sorted(meas)]} s; 128-latent window "
              f"estimated {sec_128:.1f} s (measured if 128 is listed, else two-point fit a + b*FLOP); x49 windows x50 steps; VAE omitted")
    print(json.dumps({
        "impl": "reference", "metric": "megapixels_per_sec", "value": mps, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1024x1024->4096x4096, 50 EDM steps, tiled sampler 49 windows (extrapolated from a bounded CPU sample)",
                   "sample_latent_side": side, "sec_per_sample_call": sec},
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def build_model(device):
    from supir_b200 import denoiser as dn, nets, sampling, vae, wrappers
    with torch.device(device):
        unet = nets.LightGLVUNet(mode="XL-base", project_type="ZeroSFT", project_channel_scale=2, **UNET_CFG).to(torch.bfloat16)
        ctrl = nets.GLVControl(input_upscale=1, **UNET_CFG).to(torch.bfloat16)
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=VAE_CFG, lossconfig={"target": "torch.nn.Identity"})
    net = wrappers.ControlWrapper(unet, dtype=torch.bfloat16)
    net.load_control_model(ctrl)
    net.pack()
    den = dn.DiscreteDenoiserWithControl(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config=DISC).to(device)
    smp = sampling.TiledRestoreEDMSampler(
        tile_size=TILE, tile_stride=STRIDE, tile_batch=int(os.environ.get("SUPIR_BENCH_TILE_BATCH", "49")), num_steps=EDM_STEPS,
        restore_cfg=-1.0, s_churn=5, s_noise=1.01, discretization_config=DISC,
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 1.0, "scale_min": 4.0}},
        device=device)
    ae.encoder.forward = vae.VAEHook(ae.encoder, 1024, is_decoder=False)
    ae.decoder.forward = vae.VAEHook(ae.decoder, 128, is_decoder=True)
    return net, den, smp, ae


def profile_dominant_kernel(net, device):
    """Live roofline of the dominant kernel (the tcgen05 GEMM / implicit-GEMM conv): one eager denoiser call on a
    128-latent window with a CUDA-event pair (current stream) around every GEMM-class launch; achieved = sum of the
    launches' algorithmic FLOPs / sum of their durations."""
    from supir_b200 import ops, wrappers
    rec = []
    orig_gemm, orig_conv = ops.gemm, ops.conv3x3

    def gemm(a, w, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_gemm(a, w, out, **kw)
        e1.record()
        rec.append((2.0 * a.shape[0] * w.shape[0] * a.shape[1], e0, e1, ("gemm", a.shape[0], w.shape[0], a.shape[1], kw.get("act", 0))))
        return r

    def conv(x, B, H, W, wp, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_conv(x, B, H, W, wp, out, **kw)
        e1.record()
        rec.append((2.0 * x.shape[0] * wp.shape[0] * wp.shape[1], e0, e1, ("conv3x3", x.shape[0], wp.shape[0], wp.shape[1], kw.get("act", 0))))
        return r

    # re-run, eagerly, the plan the timed steps replayed (largest batch = all windows of this rank, CFG pair)
    if net._plans:
        plan = net._plans[max(net._plans, key=lambda k: k[0])]
    else:
        B = 2 * 7
        x = torch.randn(B, 4, TILE, TILE, device=device)
        c = {"control": torch.randn(B, 4, TILE, TILE, device=device), "crossattn": torch.randn(B, 77, 2048, device=device),
             "vector": torch.randn(B, 2816, device=device)}
        t = torch.full((B,), 500, device=device)
        plan = wrappers._Plan(net, B, TILE, TILE, 77, 2048, 2816, device)
        net._fill(plan, x, t, c["crossattn"], c["vector"], c["control"], 1.0)
    plan._run()
    torch.cuda.synchronize()
    ops.gemm, ops.conv3x3 = gemm, conv
    try:
        plan._run()
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.conv3x3 = orig_gemm, orig_conv
    flops = sum(r[0] for r in rec)
    ms = sum(r[1].elapsed_time(r[2]) for r in rec)
    dump = os.environ.get("SUPIR_BENCH_DUMP_SHAPES")
    if dump:
        agg = {}
        for f, e0, e1, key in rec:
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += f
        rows = sorted(({"op": k[0], "M": k[1], "N": k[2], "K": k[3], "act": k[4], "count": v[0], "ms": round(v[1], 3),
                        "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)} for k, v in agg.items()), key=lambda r: -r["ms"])
        with open(dump, "w") as f:
            json.dump({"total_ms": ms, "total_tflop": flops / 1e12, "rows": rows, "order": [list(r[3]) for r in rec]}, f, indent=0)
    return flops, ms, len(rec)


def run_supir(args):
    import torch.distributed as dist
    from supir_b200 import _native, sampling
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    torch.manual_seed(1234)                       # same seed on every rank: identical noise draws (see sampling.py)
    net, den, smp, ae = build_model(device)
    denoiser = sampling.FusedDenoiser(den, net)
    sust, burst, hbm, peak_kind = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: pinned host buffers (synthetic LQ image already resized to the output resolution, like PIL2Tensor) ----
    img_host = torch.empty(1, 3, OUT_PX, OUT_PX, dtype=torch.float32).uniform_(-1, 1).pin_memory()
    out_host = torch.empty(1, 3, OUT_PX, OUT_PX, dtype=torch.float32).pin_memory()
    c = {"crossattn": torch.randn(1, 77, 2048, device=device), "vector": torch.randn(1, 2816, device=device)}
    uc = {"crossattn": torch.randn(1, 77, 2048, device=device), "vector": torch.randn(1, 2816, device=device)}
    scale = 0.13025
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- VAE passes before sampling (SUPIR_model.py:117-119), timed; e2e includes the image upload ----
    skip_vae = os.environ.get("SUPIR_BENCH_SKIP_VAE", "0") == "1"     # profiling aid only: invalid as a benchmark number
    from supir_b200.vae import DiagonalGaussianDistribution
    if not skip_vae:
        # untimed warm-up of the VAE passes (kernel module loads, scratch pools), the counterpart of the W warm-up EDM steps
        for _ in range(min(args.warmup, 1)):
            _w = img_host.to(device)
            _wz = scale * DiagonalGaussianDistribution(ae.quant_conv(ae.encoder(_w))).mode()
            _wx = ae.decode(1.0 / scale * _wz)
            del _w, _wz, _wx
    barrier()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    img = img_host.to(device, non_blocking=True)
    e1.record()
    if skip_vae:
        _z = 0.5 * torch.randn(1, 4, LATENT, LATENT, device=device)
        z_stage1 = 0.5 * torch.randn(1, 4, LATENT, LATENT, device=device)
    else:
        _z = scale * DiagonalGaussianDistribution(ae.quant_conv(ae.encoder(img))).mode()
        x_stage1 = ae.decode(1.0 / scale * _z)
        z_stage1 = scale * ae.encode(x_stage1)
        del x_stage1
    e2.record()
    barrier()
    h2d_img_ms, vae_pre_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    del img
    cond, ucond = dict(c, control=_z), dict(uc, control=_z)
    noised = torch.randn_like(_z)
    run = smp.begin(denoiser, noised, cond, ucond, x_center=z_stage1, control_scale=1.0)

    # ---- warm-up steps (graph capture happens in the first one) ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(args.warmup):
        run.step(i)
    barrier()
    launches_per_step = None
    # ---- timed region: EXACTLY K EDM steps (device-resident) ----
    clocks.mark()
    lc0 = _native.launch_count() + net.replayed_launches
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        run.step(i % EDM_STEPS)
    t1.record()
    barrier()
    step_ms = t0.elapsed_time(t1) / args.steps
    lc1 = _native.launch_count() + net.replayed_launches
    clk = clocks.stop() if rank == 0 else None
    # ---- e2e steps: host latent in, host latent out, every step ----
    x_host = run.x.detach().cpu().pin_memory()
    res_host = torch.empty_like(x_host).pin_memory()
    barrier()
    t2, t3 = ev(), ev()
    t2.record()
    for i in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
        run.x.copy_(x_host, non_blocking=True)
        run.step(i % EDM_STEPS)
        res_host.copy_(run.x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t3.record()
    barrier()
    e2e_step_ms = t2.elapsed_time(t3) / args.steps
    # ---- final decode (SUPIR_model.py:131) + image download ----
    e3, e4, e5 = ev(), ev(), ev()
    e3.record()
    samples = ae.decode(1.0 / scale * run.x) if not skip_vae else torch.zeros(1, 3, 8, 8, device=device)
    e4.record()
    if not skip_vae:
        out_host.copy_(samples, non_blocking=True)
    e5.record()
    barrier()
    vae_post_ms, d2h_img_ms = e3.elapsed_time(e4), e4.elapsed_time(e5)
    finite = bool(torch.isfinite(samples).all())

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_ms, e2e_step_ms = maxr(step_ms), maxr(e2e_step_ms)
    vae_ms = maxr(vae_pre_ms) + maxr(vae_post_ms)
    xfer_ms = maxr(h2d_img_ms) + maxr(d2h_img_ms)
    total_s = (EDM_STEPS * step_ms + vae_ms) / 1e3
    e2e_total_s = (EDM_STEPS * e2e_step_ms + vae_ms + xfer_ms) / 1e3
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only) ----
    roof, cpu = None, None
    g_flops, g_ms, g_n = profile_dominant_kernel(net, device)
    achieved = g_flops / (g_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "supir::gemm_tcgen05_kernel (Linear / conv1x1 / implicit-GEMM conv3x3)",
            "achieved": achieved, "peak": sust, "peak_kind": f"bf16 dense sustained, {peak_kind}", "unit": "TFLOP/s",
            "frac": achieved / sust, "traffic": load_traffic(), "traffic_unit": "bytes/launch (mean of the ncu --set full capture in profiles/)",
            "launches_timed": g_n,
            "step_flops": FLOP_WINDOW * 49, "step_achieved_tflops": FLOP_WINDOW * 49 / (step_ms * 1e-3) / 1e12 * (1.0),
            "step_frac_of_peak": FLOP_WINDOW * 49 / (step_ms * 1e-3) / 1e12 / (sust * world)}
    if world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, int(os.environ.get("SUPIR_BENCH_CPU_THREADS", "64")))
        torch.set_num_threads(cores)
        sd = oracle_state_dict()
        meas, sec_128 = probe_oracle(float(os.environ.get("SUPIR_BENCH_CPU_BUDGET_S", "25")), sd)
        mps, _ = mp_per_s_from_window_time(sec_128)
        cpu = {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": f"oracle calls (fp32, {cores} threads) of control+UNet on one window (CFG pair) at latent sides {sorted(meas)}: "
                         f"{[round(meas[k], 2) for k in sorted(meas)]} s; 128-latent window estimated {sec_128:.1f} s "
                         f"(two-point fit a + b*FLOP unless 128 was measured); x49 windows x50 steps; VAE omitted"}
    line = {
        "metric": "megapixels_per_sec", "value": MEGAPIXELS / total_s, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "1024x1024->4096x4096 (16.78 MP), 50 EDM steps, TiledRestoreEDMSampler 128/64 = 49 windows, CFG pair, "
                               "tiled VAE enc 1024 px / dec 128 latent; SUPIR-v0 + SDXL-base + SDXL-VAE shapes, random weights",
                   "edm_steps": EDM_STEPS, "windows": 49, "tile_batch": smp.tile_batch, "vae_ms": vae_ms, "vae_pre_ms": vae_pre_ms,
                   "vae_post_ms": vae_post_ms, "vae_warmup_passes": min(args.warmup, 1), "l2": "per-step working set (7.7 GB of weights + activations) exceeds the 126 MB L2",
                   "output_finite": finite, "vae_skipped_INVALID_FOR_BENCH": skip_vae, "parallelism": f"windows sharded over {world} rank(s), 1 all-gather/step" if world > 1 else "single GPU"},
        "ms_per_edm_step": step_ms,
        "e2e": {"value": MEGAPIXELS / e2e_total_s, "unit": "MP/s", "ms_per_step": e2e_step_ms,
                "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(res_host.numel() * 4),
                "image_h2d_bytes": int(img_host.numel() * 4), "image_d2h_bytes": int(out_host.numel() * 4)},
        "gpu_launches": int(lc1 - lc0),
        "clocks": clk,
        "roofline": roof,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="supir_b200", choices=["supir_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "supir_b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the supir_b200 arm has no CPU fallback (use --impl reference for the CPU oracle)")
        run_supir(args)


if __name__ == "__main__":
    main()
