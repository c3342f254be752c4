"""GPU parity of the text conditioner, through the C ABI: the four textenc.cu kernels against fp32 torch, the
product's GeneralConditionerWithControl against the REFERENCE's golden outputs (tests/golden/conditioner.npz, reduced towers) and
the two towers at their real widths against the oracle (CLIP-L: all 12 blocks of the openai/clip-vit-large-patch14 text config,
hidden state 11; bigG: width 1280 / 20 heads / MLP 5120 with a reduced depth of 6 blocks so the CPU oracle stays in seconds).
Tolerance: bf16 GEMM operands with fp32 accumulation and an fp32 residual stream against the fp32 reference -> rel. Frobenius
1.5e-2 per output (the reference itself runs these matmuls in fp16 under autocast)."""
import pytest
import torch
import torch.nn.functional as F

from conditioner_util import build_product_conditioner, check_against_golden, rel_fro
from oracle import textenc as otext
from weights import cond_batches, make_state_dict, randn

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
BF = torch.bfloat16


def test_gather_rows_and_layernorm_f32():
    from supir_b200 import ops
    table, pos = randn((1000, 768), 1).cuda(), randn((77, 768), 2).cuda()
    idx = torch.randint(0, 1000, (3 * 77,), generator=torch.Generator().manual_seed(3)).to(torch.int32)
    idx[5], idx[6] = -4, 5000                                   # clamped, like a bounds-checked nn.Embedding would refuse
    out = torch.empty(3 * 77, 768, device="cuda")
    ops.gather_rows_f32(table, idx.cuda(), out, pos=pos, L=77)
    ref = table[idx.long().clamp(0, 999).cuda()] + pos[torch.arange(3 * 77, device="cuda") % 77]
    assert torch.equal(out, ref)
    sub = torch.empty(4, 768, device="cuda")
    ops.gather_rows_f32(out, torch.tensor([0, 76, 77, 230], dtype=torch.int32, device="cuda"), sub)
    assert torch.equal(sub, out[[0, 76, 77, 230]])
    for C in (768, 1280, 192):
        x = (randn((154, C), 4) * 3 + 0.5).cuda()
        g, b = (1 + 0.1 * randn((C,), 5)).cuda(), (0.1 * randn((C,), 6)).cuda()
        yb, yf = torch.empty(154, C, dtype=BF, device="cuda"), torch.empty(154, C, device="cuda")
        ops.layernorm_f32(x, g, b, 1e-5, out_bf16=yb, out_f32=yf)
        ref = F.layer_norm(x, (C,), g, b, 1e-5)
        assert torch.allclose(yf, ref, atol=2e-5, rtol=1e-5), float((yf - ref).abs().max())
        assert torch.equal(yb, yf.to(BF))


@pytest.mark.parametrize("B,H,L", [(2, 12, 77), (3, 20, 77), (1, 2, 128), (2, 3, 1), (2, 2, 33)])
@pytest.mark.parametrize("causal", [True, False])
def test_attention_small(B, H, L, causal):
    from supir_b200 import ops
    W = H * 64
    qkv = (randn((B * L, 3 * W), 10 + L) * 1.5).to(BF).cuda()          # slices of one fused projection, like the towers use it
    out = torch.full((B * L, W), float("nan"), dtype=BF, device="cuda")
    ops.attention_small(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], out, B, H, L, causal=causal)
    sp = lambda t: t.float().reshape(B, L, H, 64).transpose(1, 2)  # noqa: E731
    ref = F.scaled_dot_product_attention(sp(qkv[:, :W]), sp(qkv[:, W:2 * W]), sp(qkv[:, 2 * W:]), is_causal=causal)
    ref = ref.transpose(1, 2).reshape(B * L, W)
    assert torch.isfinite(out.float()).all()
    assert torch.allclose(out.float(), ref, atol=1e-2, rtol=1e-2), float((out.float() - ref).abs().max())


def test_activation_modes():
    from supir_b200 import ops
    x = (randn((154, 5120), 20) * 2).to(BF).cuda()
    for mode, fn in (("gelu", F.gelu), ("quick_gelu", lambda t: t * torch.sigmoid(1.702 * t))):
        y = torch.empty_like(x)
        ops.activation(x, y, mode)
        assert torch.allclose(y.float(), fn(x.float()), atol=1e-2, rtol=1e-2), mode
        z = x.clone()
        ops.activation(z, z, mode)                                # in place
        assert torch.equal(z, y)


def test_conditioner_vs_reference_golden():
    from supir_b200 import _native
    batch, batch_uc, tl, tg = cond_batches()
    gc = build_product_conditioner(tl, tg, device="cuda")
    to_cuda = lambda b: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}  # noqa: E731
    _native.reset_launch_count()
    b = to_cuda(batch)
    worst = check_against_golden(gc, b, to_cuda(batch_uc), tol=1.5e-2)
    assert _native.launch_count() > 100, "CUDA kernels were not launched"
    print(f"conditioner vs reference golden: worst rel. Frobenius {worst:.3g}")


def _tokens(n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(1, 49405, (n, 77), generator=g)
    t[:, 0] = 49406
    for i in range(n):
        k = int(torch.randint(4, 76, (1,), generator=g))
        t[i, k:] = 49407
    return t


def test_clip_l_full_size_vs_oracle():
    from supir_b200 import conditioner as C
    e = C.FrozenCLIPEmbedder(layer="hidden", layer_idx=11, always_return_pooled=True)
    sd = make_state_dict({k: list(v.shape) for k, v in e.state_dict().items()}, seed=5)
    e.load_state_dict(sd)
    e = e.cuda()
    tok = _tokens(3, 6)
    z, pooled = e(tok.cuda())
    rz, rp = otext.frozen_clip_embedder(sd, tok, 12, "hidden", 11, True)
    ez, ep = rel_fro(z.cpu(), rz), rel_fro(pooled.cpu(), rp)
    print(f"CLIP-L (12 blocks, 768 wide): hidden[11] rel. Frobenius {ez:.3g}, pooled {ep:.3g}")
    assert z.shape == (3, 77, 768) and ez <= 1.5e-2 and ep <= 1.5e-2


def test_bigg_full_width_vs_oracle():
    from supir_b200 import conditioner as C
    e = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", layer="penultimate", always_return_pooled=True, legacy=False, text_cfg={"layers": 6})
    sd = make_state_dict({k: list(v.shape) for k, v in e.state_dict().items()}, seed=7)
    e.load_state_dict(sd)
    e = e.cuda()
    tok = _tokens(2, 8)
    tok[:, :][tok == 49407] = 0                                     # open_clip pads with 0 ...
    for i in range(2):
        tok[i, int((tok[i] == 0).nonzero()[0])] = 49407            # ... after ONE end-of-text token
    z, pooled = e(tok.cuda())
    rz, rp = otext.frozen_openclip_embedder2(sd, tok, 20, "penultimate", True, False)
    ez, ep = rel_fro(z.cpu(), rz), rel_fro(pooled.cpu(), rp)
    print(f"bigG (6 blocks, 1280 wide): penultimate rel. Frobenius {ez:.3g}, pooled {ep:.3g}")
    assert z.shape == (2, 77, 1280) and pooled.shape == (2, 1280) and ez <= 1.5e-2 and ep <= 2e-2
