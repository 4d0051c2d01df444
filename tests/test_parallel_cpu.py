"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): tile sharding + halo exchange of the tiled
sampler must equal the single-process blend BIT FOR BIT (contributions are added in the global window
order on every rank), for one image and for a pooled (image, window) work list; image sharding must
cover every image exactly once."""
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


def _fake_step(x_tile, i, noise_tile, win, img=0):
    """Deterministic stand-in for one denoiser + Euler update of a window (depends on the window position and
    the image so ownership mistakes show up)."""
    h0, _, w0, _ = win
    return 0.9 * x_tile + 0.1 * noise_tile + 0.001 * (h0 + 2 * w0 + 5 * img) + 0.01 * (i + 1) * torch.tanh(x_tile)


def _reference(x, noises, H, W, tile, stride, steps, img=0):
    """sampling.py:716-756 in one process: acc[win] += tile * w in window order, then / count."""
    from b200sr.sampling import gaussian_weights, sliding_windows

    wgt = gaussian_weights(tile, tile)
    for i in range(steps):
        acc, cnt = torch.zeros_like(x), torch.zeros_like(x)
        for win in sliding_windows(H, W, tile, stride):
            h0, h1, w0, w1 = win
            acc[:, :, h0:h1, w0:w1] += _fake_step(x[:, :, h0:h1, w0:w1], i, noises[i][:, :, h0:h1, w0:w1], win, img) * wgt
            cnt[:, :, h0:h1, w0:w1] += wgt
        x = acc / cnt
    return x


def _inputs(n_images, H, W, steps):
    g = torch.Generator().manual_seed(7)          # identical on every rank
    xs = {m: torch.randn(1, 4, H, W, generator=g) for m in range(n_images)}
    noises = {m: [torch.randn(1, 4, H, W, generator=g) for _ in range(steps)] for m in range(n_images)}
    return xs, noises


def _worker(rank, world, port, n_images, H, W, tile, stride, steps, tile_batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200sr.parallel import PooledTileStepper, TileShardedStepper

        xs, noises = _inputs(n_images, H, W, steps)
        if n_images == 1 and tile_batch == 1:      # the single-image wrapper
            st = TileShardedStepper(H, W, tile, stride)
            x = xs[0]
            for i in range(steps):
                x = st.step(x, i, noises[0][i], _fake_step)
            fulls = [st.gather_full(x)]
        else:
            st = PooledTileStepper(n_images, H, W, tile, stride, tile_batch=tile_batch)

            def step_fn(m, wins, x_tiles, i, noise_tiles):
                outs = [_fake_step(x_tiles[j:j + 1], i, noise_tiles[j:j + 1], w, m) for j, w in enumerate(wins)]
                return torch.cat(outs, 0)

            cur = {m: xs[m] for m in st.my_images}
            for i in range(steps):
                cur = st.step(cur, i, {m: noises[m][i] for m in st.my_images}, step_fn)
            fulls = [st.gather_image(m, cur.get(m), xs[0]) for m in range(n_images)]
        if rank == 0:
            q.put((fulls, st.halo_bytes_per_step, [len(p) for p in st.parts], sorted(st.strips.keys())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_images,H,W,tile,stride,tile_batch", [
    (2, 1, 64, 64, 32, 24, 1), (3, 1, 64, 48, 32, 16, 1), (2, 1, 32, 32, 32, 24, 1),
    (3, 2, 64, 64, 32, 24, 1),      # 18 units over 3 ranks: image 0 and 1 both straddle a rank boundary
    (2, 3, 64, 64, 32, 24, 2),      # pooled list with two windows per network call
])
def test_tile_sharding_equals_single_process_bitwise(world, n_images, H, W, tile, stride, tile_batch):
    steps = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, H, W, tile, stride, steps, tile_batch, q))
             for r in range(world)]
    for p in procs:
        p.start()
    fulls, halo_bytes, counts, pairs = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xs, noises = _inputs(n_images, H, W, steps)
    for m in range(n_images):
        ref = _reference(xs[m], noises[m], H, W, tile, stride, steps, m)
        assert torch.equal(fulls[m], ref), f"image {m}: max diff {(fulls[m] - ref).abs().max().item():.3e}"
    from b200sr.sampling import sliding_windows

    nw = len(sliding_windows(H, W, tile, stride))
    assert sum(counts) == n_images * nw
    if nw > 1:
        assert 0 < halo_bytes < 4 * 4 * H * W * n_images  # only strips travel, never the whole latent
        assert all(a != b for a, b in pairs)


def test_partition_cover_and_plan():
    from b200sr.parallel import PooledTileStepper, disjoint_cover, partition_windows, shard_images
    from b200sr.sampling import sliding_windows

    wins = sliding_windows(256, 256, 128, 96)      # BASELINE config 4: 9 windows
    assert len(wins) == 9
    for world in (1, 2, 4, 8, 16):
        parts = partition_windows(wins, world)
        assert len(parts) == world and sorted(w for p in parts for w in p) == sorted(wins)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # disjoint_cover: same union, no pixel twice
    rects = [(0, 32, 0, 128), (0, 128, 96, 128), (16, 48, 64, 112), None]
    cover = disjoint_cover(rects)
    a, b = torch.zeros(128, 128), torch.zeros(128, 128)
    for r in rects:
        if r is not None:
            a[r[0]:r[1], r[2]:r[3]] = 1
    for r in cover:
        b[r[0]:r[1], r[2]:r[3]] += 1
    assert torch.equal(a, b)
    # single process: the stepper's plan is empty and the count matches the windows' weights
    st = PooledTileStepper(10, 256, 256, 128, 96)
    assert len(st.units) == 90 and st.strips == {} and st.halo_bytes_per_step == 0
    assert float(st.count.min()) > 0
    got = sorted(i for r in range(8) for i in shard_images(64, r, 8))
    assert got == list(range(64)) and all(len(shard_images(64, r, 8)) == 8 for r in range(8))
