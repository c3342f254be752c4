"""Bandwidth of the HBM-bound kernels (LayerNorm, GroupNorm statistics / apply) at the shapes of the 49-window step
(batch 98): achieved GB/s on buffers rotated so that no launch finds its input in L2."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supir_b200 import ops

dev = "cuda"
B = int(os.environ.get("BENCH_B", "98"))


def timeit(fn, n_rot, iters=12):
    for i in range(n_rot):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rot = 3
    for rows, C in ((B * 1024, 1280), (B * 4096, 640)):
        xs = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(rot)]
        ys = [torch.empty_like(x) for x in xs]
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ms = timeit(lambda i: ops.layernorm(xs[i], ys[i], g, b), rot)
        print(json.dumps({"op": "layernorm", "rows": rows, "C": C, "ms": round(ms, 4), "GBps": round(2 * rows * C * 2 / ms / 1e6, 1)}), flush=True)
        del xs, ys
    for HW, C in ((16384, 320), (4096, 640), (1024, 1280), (16384, 640), (4096, 1280), (1024, 2560)):
        rows = B * HW
        xs = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(rot)]
        ys = [torch.empty_like(x) for x in xs]
        ws = torch.empty(ops.groupnorm_ws_size(B, HW, C), dtype=torch.float64, device=dev)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ms = timeit(lambda i: ops.groupnorm_stats(xs[i], B, HW, ws), rot)
        print(json.dumps({"op": "gn_stats", "HW": HW, "C": C, "ms": round(ms, 4), "GBps": round(rows * C * 2 / ms / 1e6, 1)}), flush=True)
        ms = timeit(lambda i: ops.groupnorm_apply(xs[i], B, HW, ys[i], g, b, 1e-5, True, sums=ws), rot)
        print(json.dumps({"op": "gn_apply", "HW": HW, "C": C, "ms": round(ms, 4), "GBps": round(2 * rows * C * 2 / ms / 1e6, 1)}), flush=True)
        del xs, ys


main()
