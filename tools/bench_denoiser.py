"""Quick timing of one denoiser network call (GLVControl + LightGLVUNet, full SDXL depth) on random weights."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supir_b200 import nets, wrappers, _native

CFG = dict(adm_in_channels=2816, num_classes="sequential", use_checkpoint=True, in_channels=4, out_channels=4,
           model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4], num_head_channels=64,
           use_spatial_transformer=True, use_linear_in_transformer=True, transformer_depth=[1, 2, 10], context_dim=2048,
           spatial_transformer_attn_type="softmax-xformers", legacy=False)
FLOP = {64: 4.766e12, 128: 20.292e12, 256: 107.887e12}

def record_shapes(path):
    """Wrap ops.gemm / ops.conv3x3 to log the launch order of GEMM-class kernels (joined offline with an ncu launch list)."""
    from supir_b200 import ops
    log = []
    og, oc = ops.gemm, ops.conv3x3

    def gemm(a, w, out, **kw):
        log.append(["gemm", a.shape[0], w.shape[0], a.shape[1], kw.get("act", 0), kw.get("residual") is not None])
        return og(a, w, out, **kw)

    def conv(x, B, H, W, wp, out, **kw):
        log.append(["conv3x3", x.shape[0], wp.shape[0], wp.shape[1], kw.get("act", 0), kw.get("residual") is not None])
        return oc(x, B, H, W, wp, out, **kw)
    ops.gemm, ops.conv3x3 = gemm, conv
    import atexit
    atexit.register(lambda: json.dump(log, open(path, "w")))


def main():
    if os.environ.get("BENCH_SHAPE_LOG"):
        record_shapes(os.environ["BENCH_SHAPE_LOG"])
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128]
    t0 = time.time()
    with torch.device("cuda"):
        unet = nets.LightGLVUNet(mode="XL-base", project_type="ZeroSFT", project_channel_scale=2, **CFG).to(torch.bfloat16)
        ctrl = nets.GLVControl(input_upscale=1, **CFG).to(torch.bfloat16)
    w = wrappers.ControlWrapper(unet, dtype=torch.bfloat16)
    w.load_control_model(ctrl)
    w.pack()
    torch.cuda.synchronize()
    print(f"build+pack {time.time()-t0:.1f}s, mem {torch.cuda.memory_allocated()/2**30:.1f} GiB", flush=True)
    for hw in sizes:
        for B in [int(b) for b in os.environ.get('BENCH_B', '2,4,8').split(',')]:
            if hw >= 256 and B > 2:
                continue
            x = torch.randn(B, 4, hw, hw, device="cuda")
            c = {"control": torch.randn(B, 4, hw, hw, device="cuda"), "crossattn": torch.randn(B, 77, 2048, device="cuda"),
                 "vector": torch.randn(B, 2816, device="cuda")}
            t = torch.full((B,), 500, device="cuda")
            _native.reset_launch_count()
            out = w(x, t, c, 1.0)
            torch.cuda.synchronize()
            launches = _native.launch_count() // 2   # warm-up + capture
            for _ in range(2):
                w(x, t, c, 1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 5
            e0.record()
            for _ in range(iters):
                w(x, t, c, 1.0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = FLOP.get(hw, 0) * B / 2
            print(json.dumps({"latent": hw, "B": B, "ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1), "launches": launches,
                              "finite": bool(torch.isfinite(out).all()), "mem_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)

main()
