"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): tile sharding + halo exchange of the tiled
sampler must equal the single-process blend; image sharding must cover every image exactly once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_step(x_tile, i, noise_tile, win):
    """Deterministic stand-in for one denoiser + Euler update of a window (depends on the window position
    so ownership mistakes show up)."""
    h0, _, w0, _ = win
    return 0.9 * x_tile + 0.1 * noise_tile + 0.001 * (h0 + 2 * w0) + 0.01 * (i + 1) * torch.tanh(x_tile)


def _accumulate(tile, weight, acc, h0, w0):
    th, tw = tile.shape[-2:]
    acc[:, :, h0:h0 + th, w0:w0 + tw] += tile * weight


def _reference(x, noises, H, W, tile, stride, steps):
    from b200sr.sampling import gaussian_weights, sliding_windows

    wgt = gaussian_weights(tile, tile)
    for i in range(steps):
        acc, cnt = torch.zeros_like(x), torch.zeros_like(x)
        for win in sliding_windows(H, W, tile, stride):
            h0, h1, w0, w1 = win
            acc[:, :, h0:h1, w0:w1] += _fake_step(x[:, :, h0:h1, w0:w1], i, noises[i][:, :, h0:h1, w0:w1], win) * wgt
            cnt[:, :, h0:h1, w0:w1] += wgt
        x = acc / cnt
    return x


def _worker(rank, world, port, H, W, tile, stride, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200sr.parallel import TileShardedStepper

        g = torch.Generator().manual_seed(7)          # identical on every rank
        x = torch.randn(1, 4, H, W, generator=g)
        noises = [torch.randn(1, 4, H, W, generator=g) for _ in range(steps)]
        st = TileShardedStepper(H, W, tile, stride)
        for i in range(steps):
            x = st.step(x, i, noises[i], _fake_step, _accumulate)
        full = st.gather_full(x)
        if rank == 0:
            q.put((full, st.halo_bytes_per_step, [len(p) for p in st.parts], sorted(st.plan.keys())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,tile,stride", [(2, 64, 64, 32, 24), (3, 64, 48, 32, 16), (2, 32, 32, 32, 24)])
def test_tile_sharding_equals_single_process(world, H, W, tile, stride):
    steps = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, tile, stride, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, halo_bytes, counts, pairs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 4, H, W, generator=g)
    noises = [torch.randn(1, 4, H, W, generator=g) for _ in range(steps)]
    ref = _reference(x, noises, H, W, tile, stride, steps)
    assert torch.allclose(full, ref, rtol=1e-5, atol=1e-6)
    from b200sr.sampling import sliding_windows

    assert sum(counts) == len(sliding_windows(H, W, tile, stride))
    if len(sliding_windows(H, W, tile, stride)) > 1:
        assert halo_bytes > 0 and halo_bytes < 4 * 4 * H * W  # only strips travel, never the whole latent
        assert all(a != b for a, b in pairs)


def test_partition_and_plan():
    from b200sr.parallel import halo_plan, partition_windows, shard_images
    from b200sr.sampling import sliding_windows

    wins = sliding_windows(256, 256, 128, 96)      # BASELINE config 4: 9 windows
    assert len(wins) == 9
    for world in (1, 2, 4, 8, 16):
        parts = partition_windows(wins, world)
        assert len(parts) == world and sorted(w for p in parts for w in p) == sorted(wins)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        plan = halo_plan(parts)
        for (s, d), rects in plan.items():
            assert (d, s) in plan and len(rects) == len(plan[(d, s)])
            for (h0, h1, w0, w1) in rects:
                assert 0 <= h0 < h1 <= 256 and 0 <= w0 < w1 <= 256
    got = sorted(i for r in range(8) for i in shard_images(64, r, 8))
    assert got == list(range(64)) and all(len(shard_images(64, r, 8)) == 8 for r in range(8))
