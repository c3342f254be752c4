"""Times the text conditioner of options/SUPIR_v0.yaml at its REAL size (CLIP-L: 12 blocks x 768, hidden state 11; OpenCLIP bigG:
32 blocks x 1280, penultimate + pooled; three size embedders) for one prompt pair (c and uc, as SUPIRModel.prepare_condition
calls it), with CUDA events, beside the same towers as plain PyTorch (the oracle's functional restatement on cuda under bf16
autocast = what transformers / open_clip execute on a GPU). Random weights (none exist offline). Measuring tool, not product.

    python tools/bench_conditioner.py [--small]     # --small: reduced towers, for a CPU dry run of the script's logic
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tokens(n, pad, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.full((n, 77), pad, dtype=torch.long)
    for i in range(n):
        k = int(torch.randint(5, 60, (1,), generator=g))
        t[i, 0], t[i, 1:1 + k], t[i, 1 + k] = 49406, torch.randint(1, 49405, (k,), generator=g), 49407
    return t


def main():
    small = "--small" in sys.argv
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    if dev == "cpu":
        import cpu_ops
        cpu_ops.install(None)
    from oracle import textenc as otext
    from supir_b200 import _native
    from supir_b200.config import instantiate_from_config
    la, ga = ({"layers": 2, "vocab": 49408}, {"layers": 2}) if small else ({}, {})
    nl, ng = (2, 2) if small else (12, 32)
    emb = [{"input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenCLIPEmbedder", "params": {"layer": "hidden", "layer_idx": nl - 1, "arch": la}},
           {"input_key": "txt", "target": "sgm.modules.encoders.modules.FrozenOpenCLIPEmbedder2",
            "params": {"arch": "ViT-bigG-14", "layer": "penultimate", "always_return_pooled": True, "legacy": False, "text_cfg": ga}}]
    emb += [{"input_key": k, "target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 256}}
            for k in ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple")]
    torch.manual_seed(0)
    with torch.device(dev):
        gc = instantiate_from_config({"target": "sgm.modules.GeneralConditionerWithControl", "params": {"emb_models": emb}})
    for p in gc.parameters():                      # LayerNorm scales stay 1, everything else small random: finite activations
        if p.dim() >= 2:
            p.data.normal_(0, 0.02)
    tl, tg = tokens(1, 49407, 1), tokens(1, 0, 2)
    tlu, tgu = tokens(1, 49407, 3), tokens(1, 0, 4)
    table = {"p": (tl, tg), "n": (tlu, tgu)}
    gc.embedders[0].tokenize = lambda texts: torch.cat([table[t][0] for t in texts])
    gc.embedders[1].tokenize = lambda texts: torch.cat([table[t][1] for t in texts])
    size = lambda v: torch.tensor([v], device=dev)  # noqa: E731
    batch = {"txt": ["p"], "control": torch.zeros(1, 4, 8, 8, device=dev), "original_size_as_tuple": size([1024, 1024]),
             "crop_coords_top_left": size([0, 0]), "target_size_as_tuple": size([1024, 1024])}
    batch_uc = dict(batch, txt=["n"])
    sync = torch.cuda.synchronize if dev == "cuda" else (lambda: None)

    def timed(fn, reps):
        fn()
        sync()
        if dev == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            sync()
            return e0.elapsed_time(e1) / reps
        t0 = time.time()
        for _ in range(reps):
            fn()
        return (time.time() - t0) * 1e3 / reps

    out = {}
    ours = lambda: gc.get_unconditional_conditioning(dict(batch), dict(batch_uc))  # noqa: E731
    c, uc = ours()
    _native.reset_launch_count() if dev == "cuda" else None
    ours()
    launches = _native.launch_count() if dev == "cuda" else 0
    out["supir_b200_ms"] = timed(ours, 5)
    out["launches_per_prompt_pair"] = launches
    # the same arithmetic as plain PyTorch on the device (bf16 autocast, fp32 weights like the reference keeps them)
    sd = {k: v.detach() for k, v in gc.state_dict().items()}

    def torch_path():
        res = []
        for (a, b), bt in (((tl, tg), batch), ((tlu, tgu), batch_uc)):
            ob = dict(bt, txt_tokens_l=a.to(dev), txt_tokens_g=b.to(dev))
            with torch.autocast(dev, dtype=torch.bfloat16):
                res.append(otext.supir_conditioner(sd, ob, 12, 20, clip_layer_idx=nl - 1))
        return res

    if dev == "cuda":
        with torch.device(dev):                     # the oracle builds its causal mask / frequencies with default-device tensors
            rc, ruc = torch_path()
            out["torch_bf16_autocast_ms"] = timed(torch_path, 5)
    else:
        rc, ruc = torch_path()
        out["torch_bf16_autocast_ms"] = timed(torch_path, 1)
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
    out["rel_fro_vs_torch_path"] = {k: max(rel(c[k], rc[k]), rel(uc[k], ruc[k])) for k in ("crossattn", "vector")}
    out["config"] = {"clip_l_blocks": nl, "bigg_blocks": ng, "prompts": "1 positive + 1 negative", "device": dev}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
