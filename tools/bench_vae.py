"""Timing of the tiled VAE passes of the bench workload (encode 4096^2 px in 1024-px tiles, decode 512^2 latent in
128-latent tiles) with a per-op breakdown: every supir_b200.ops launch is bracketed by a CUDA-event pair."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supir_b200 import ops, vae

VAE_CFG = dict(attn_type="vanilla-xformers", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
               ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
NAMES = ["conv3x3", "conv_geom", "gemm", "attention_1head", "groupnorm_stats", "groupnorm_apply", "groupnorm_finalize", "groupnorm_merge_tiles",
         "upsample2x", "axpy", "copy2d", "conv3x3_small_cin", "conv3x3_small_cout", "im2col_s2", "softmax_rows"]


def instrument(rec):
    orig = {}
    for n in NAMES:
        f = getattr(ops, n)
        orig[n] = f

        def wrap(*a, _f=f, _n=n, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = _f(*a, **kw)
            e1.record()
            shp = tuple(tuple(t.shape) for t in a[:3] if torch.is_tensor(t))
            rec.append((_n, shp, e0, e1))
            return r
        setattr(ops, n, wrap)
    return orig


def main():
    px = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = "cuda"
    with torch.device(dev):
        ae = vae.AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=VAE_CFG, lossconfig={"target": "torch.nn.Identity"})
    ae.encoder.forward = vae.VAEHook(ae.encoder, 1024, is_decoder=False)
    ae.decoder.forward = vae.VAEHook(ae.decoder, 128, is_decoder=True)
    img = torch.empty(1, 3, px, px, device=dev).uniform_(-1, 1)
    z = 0.5 * torch.randn(1, 4, px // 8, px // 8, device=dev)
    for _ in range(2):
        ae.encode(img); ae.decode(z)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e = [ev() for _ in range(3)]
    e[0].record(); ae.encode(img); e[1].record(); ae.decode(z); e[2].record()
    torch.cuda.synchronize()
    print(json.dumps({"px": px, "encode_ms": e[0].elapsed_time(e[1]), "decode_ms": e[1].elapsed_time(e[2])}), flush=True)
    if os.environ.get("VAE_BREAKDOWN", "1") == "1":
        for name, fn in (("encode", lambda: ae.encode(img)), ("decode", lambda: ae.decode(z))):
            rec = []
            orig = instrument(rec)
            try:
                e0, e1 = ev(), ev()
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
            finally:
                for n, f in orig.items():
                    setattr(ops, n, f)
            agg = collections.OrderedDict()
            for n, shp, a, b in rec:
                k = (n, shp)
                v = agg.setdefault(k, [0, 0.0])
                v[0] += 1; v[1] += a.elapsed_time(b)
            tot = sum(v[1] for v in agg.values())
            print(f"== {name}: wall {e0.elapsed_time(e1):.1f} ms, sum of op times {tot:.1f} ms, {len(rec)} launches")
            byop = collections.Counter()
            for (n, shp), (c, t) in agg.items():
                byop[n] += t
            for n, t in byop.most_common():
                print(f"   {n:24s} {t:9.2f} ms {100*t/tot:5.1f}%")
            for (n, shp), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
                print(f"   {t:8.2f} ms n={c:4d} avg={1e3*t/c:8.1f}us {n} {shp}")


main()
