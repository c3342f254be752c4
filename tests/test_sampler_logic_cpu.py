"""Host-side sampler logic on CPU: the samplers of supir_b200.sampling driven with plain-torch stand-ins for the handful of
kernels they call (tests/cpu_ops.py) reproduce the REFERENCE's golden runs (tests/golden/sampler_toy.npz: toy network,
injected noise) on the fused path — tiled (CFG branch, window) units with every tile_batch, per-window prompts, the untiled
run object — and, under gloo with 2 and 3 ranks, the sharded runs equal the single-process result bit for bit (uneven unit
split and the CFG-branch split of the untiled sampler included). The arithmetic of the real kernels is covered by the -m gpu
tests; this file pins the Python around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from weights import randn

G = os.path.join(os.path.dirname(__file__), "golden")
DISC = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.LinearCFG", "params": {"scale": 1.0, "scale_min": 4.0}}


def toy_network(x, t, c, control_scale):
    tt = (t.float() / 1000.0).view(-1, 1, 1, 1)
    v = c["vector"].mean(dim=1).view(-1, 1, 1, 1)
    return 0.3 * torch.tanh(x) + 0.1 * tt + 0.2 * control_scale * c["control"] + 0.05 * v


class SeededNoise:
    def __init__(self, base):
        self.base, self.n = base, 0

    def __call__(self, x, **k):
        self.n += 1
        return randn(tuple(x.shape), self.base + self.n).to(x.device, x.dtype)


def make_denoiser():
    from supir_b200 import denoiser as dn
    return dn.DiscreteDenoiserWithControl(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config=DISC)


def tiled_inputs():
    x = randn((1, 4, 40, 28), 60)
    c = {"control": randn((1, 4, 40, 28), 61), "vector": randn((1, 6), 62), "crossattn": randn((1, 3, 5), 63)}
    uc = {"control": c["control"], "vector": randn((1, 6), 64), "crossattn": randn((1, 3, 5), 65)}
    return x, c, uc, randn((1, 4, 40, 28), 66)


def test_fused_paths_match_reference_goldens(monkeypatch):
    import cpu_ops
    from supir_b200 import sampling
    cpu_ops.install(monkeypatch)
    g = np.load(os.path.join(G, "sampler_toy.npz"))
    denoiser = sampling.FusedDenoiser(make_denoiser(), toy_network)
    for name, restore_cfg, lin_cs in [("edm", -1.0, False), ("edm_restore", 4.0, True)]:
        smp = sampling.RestoreEDMSampler(num_steps=6, restore_cfg=restore_cfg, s_churn=5, s_noise=1.01, discretization_config=DISC,
                                         guider_config=GUIDER, device="cpu")
        x = randn((2, 4, 12, 10), 50)
        c = {"control": randn((2, 4, 12, 10), 51), "vector": randn((2, 6), 52), "crossattn": randn((2, 3, 5), 53)}
        uc = {"control": c["control"], "vector": randn((2, 6), 54), "crossattn": randn((2, 3, 5), 55)}
        monkeypatch.setattr(torch, "randn_like", SeededNoise(1000))
        out = smp(denoiser, x, cond=c, uc=uc, x_center=randn((2, 4, 12, 10), 56), control_scale=0.9, use_linear_control_scale=lin_cs,
                  control_scale_start=0.2)
        torch.testing.assert_close(out, torch.from_numpy(g[name]), rtol=1e-4, atol=1e-4)
    x, c, uc, xc = tiled_inputs()
    for tile_batch in (1, 3, 100):
        smp = sampling.TiledRestoreEDMSampler(tile_size=16, tile_stride=8, tile_batch=tile_batch, num_steps=4, restore_cfg=4.0, s_churn=5,
                                              s_noise=1.01, discretization_config=DISC, guider_config=GUIDER, device="cpu")
        monkeypatch.setattr(torch, "randn_like", SeededNoise(2000))
        out = smp(denoiser, x, cond=c, uc=uc, x_center=xc, control_scale=1.0)
        torch.testing.assert_close(out, torch.from_numpy(g["tiled"]), rtol=1e-4, atol=1e-4)
        nwin = len(sampling._sliding_windows(40, 28, 16, 8))
        conds = [{"control": c["control"], "vector": randn((1, 6), 700 + j), "crossattn": randn((1, 3, 5), 800 + j)} for j in range(nwin)]
        monkeypatch.setattr(torch, "randn_like", SeededNoise(2500))
        out = smp(denoiser, x, cond=conds, uc=uc, x_center=xc, control_scale=1.0)
        torch.testing.assert_close(out, torch.from_numpy(g["tiled_local"]), rtol=1e-4, atol=1e-4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import torch.distributed as dist
    import cpu_ops
    from supir_b200 import sampling
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu_ops.install()
    denoiser = sampling.FusedDenoiser(make_denoiser(), toy_network)
    x, c, uc, xc = tiled_inputs()

    def tiled(shard):
        smp = sampling.TiledRestoreEDMSampler(tile_size=16, tile_stride=8, tile_batch=2, num_steps=3, restore_cfg=4.0, s_churn=5,
                                              s_noise=1.01, discretization_config=DISC, guider_config=GUIDER, device="cpu")
        smp.shard = shard
        torch.randn_like = SeededNoise(2000)
        return smp(denoiser, x, cond=c, uc=uc, x_center=xc, control_scale=1.0)

    def untiled(shard):
        smp = sampling.RestoreEDMSampler(num_steps=4, restore_cfg=4.0, s_churn=5, s_noise=1.01, discretization_config=DISC,
                                         guider_config=GUIDER, device="cpu")
        smp.shard = shard
        torch.randn_like = SeededNoise(1000)
        return smp(denoiser, x, cond=c, uc=uc, x_center=xc, control_scale=0.9)

    a, b = tiled(True), tiled(False)
    ua, ub = untiled(True), untiled(False)
    q.put((rank, bool(torch.equal(a, b)), bool(torch.equal(ua, ub))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 5])        # 24 units: 12 + 12, and 5 + 5 + 5 + 5 + 4 (padded exchange + compaction)
def test_sharded_runs_equal_single_process_under_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] and r[2] for r in res), res
