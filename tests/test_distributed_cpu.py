"""world_size-2 gloo test of the multi-GPU plumbing on CPU: window sharding + the per-step all-gather reproduce the
single-process result of the reference's sequential blend bit-for-bit."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _blend_reference_order(tiles, windows, weights64, shape):
    """sampling.py:656-659 in the reference's order (the oracle of supir_tile_blend)."""
    x_next, count = torch.zeros(shape), torch.zeros(shape)
    tw = weights64.repeat(shape[0], shape[1], 1, 1)
    for j, (hi, he, wi, we) in enumerate(windows):
        if hi < 0:
            continue
        x_next[:, :, hi:he, wi:we] += tiles[j] * tw
        count[:, :, hi:he, wi:we] += tw
    x_next /= count
    return x_next


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import sampler as osamp
    from supir_b200.sampling import _sliding_windows, exchange_window_outputs, shard_windows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, C, H, W, T = 1, 4, 40, 28, 16
    windows = _sliding_windows(H, W, T, 8)
    nw = len(windows)
    per, lo, hi = shard_windows(nw, world, rank)
    g = torch.Generator().manual_seed(7)
    all_tiles = torch.randn((nw, N, C, T, T), generator=g)      # what a single process would compute
    tiles_out = torch.zeros((per * world, N, C, T, T))
    tiles_out[lo:hi] = all_tiles[lo:hi]                         # this rank computes only its windows
    exchange_window_outputs(tiles_out, rank, per)
    table = [(-1, -1, -1, -1)] * (per * world)
    table[:nw] = windows
    w64 = torch.tensor(osamp.gaussian_weights(T, T))
    got = _blend_reference_order(tiles_out, table, w64, (N, C, H, W))
    want = _blend_reference_order(all_tiles, windows, w64, (N, C, H, W))
    q.put((rank, bool(torch.equal(got, want)), bool(torch.equal(tiles_out[:nw], all_tiles))))
    dist.barrier()
    dist.destroy_process_group()


def _unit_worker(rank, world, port, q):
    """Fused path: (CFG branch, window) units in balanced contiguous ranges, padded all-gather, compaction."""
    import torch.distributed as dist
    from supir_b200.sampling import exchange_unit_outputs, shard_units
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nw, N, C, T = 7, 1, 4, 8                      # 14 units over 3 ranks -> 5, 5, 4
    g = torch.Generator().manual_seed(11)
    all_units = torch.randn((2 * nw, N, C, T, T), generator=g)
    per, lo, hi = shard_units(2 * nw, world, rank)
    pad = torch.zeros((per * world, N, C, T, T))
    pad[rank * per:rank * per + (hi - lo)] = all_units[lo:hi]
    exchange_unit_outputs(pad, rank, per)
    compact = torch.zeros_like(all_units)
    for r in range(world):
        _, l, h = shard_units(2 * nw, world, r)
        compact[l:h] = pad[r * per:r * per + (h - l)]
    q.put((rank, bool(torch.equal(compact, all_units)), hi - lo))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_units_all_gather_equals_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_unit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == 14 and max(r[2] for r in res) - min(r[2] for r in res) <= 1


def _vae_worker(rank, world, port, q):
    """Tiled-VAE output exchange: round-robin tiles, crops packed back to back per rank, one all-gather, paste."""
    import torch.distributed as dist
    from supir_b200 import vae
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for (h, w, ts, dec) in [(40, 52, 16, True), (192, 160, 64, False), (75, 33, 16, True)]:
        N, C = 1, 3
        in_b, out_b = vae.split_tiles(h, w, ts, dec)
        sc = (lambda v: v * 8) if dec else (lambda v: v // 8)
        dims = [((b[3] - b[2]) * 8, (b[1] - b[0]) * 8) if dec else ((b[3] - b[2]) // 8, (b[1] - b[0]) // 8) for b in in_b]
        crops = [vae.crop_margins(dims[i][0], dims[i][1], in_b[i], out_b[i], dec) for i in range(len(in_b))]
        canvas = torch.randn((N, C, sc(h), sc(w)), generator=torch.Generator().manual_seed(5))
        numel, offs, slot = vae.plan_packed_crops(crops, N * C, world)
        packed = torch.zeros((world, slot))
        for i in range(len(in_b)):
            if i % world == rank:
                ob = out_b[i]
                packed[rank, offs[i]:offs[i] + numel[i]] = canvas[:, :, ob[2]:ob[3], ob[0]:ob[1]].reshape(-1)
        dist.all_gather_into_tensor(packed.view(-1), packed[rank].clone())
        result = torch.full_like(canvas, float("nan"))
        vae.paste_packed_crops(result, packed, crops, out_b, numel, offs, world)
        ok = ok and bool(torch.equal(result, canvas))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_vae_cropped_tile_exchange_reassembles_the_canvas(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vae_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world)) and all(r[1] for r in res), res


@pytest.mark.parametrize("world", [2])
def test_sharded_windows_all_gather_equals_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] and r[2] for r in res), res
