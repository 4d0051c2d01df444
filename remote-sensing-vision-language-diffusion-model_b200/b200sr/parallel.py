"""Multi-GPU sharding of the denoiser path: one process per GPU (torch.distributed; NCCL over
NVLink on the B200 box, gloo in the CPU tests).

The path shards two ways and never needs model parallelism (7.7 GB of bf16
weights fit on every GPU):

* images (infer_dir.py:198-202 processes them in a sequential loop): ``shard_images`` deals them
  round-robin, no communication during sampling;
* latent tiles of one large image (TiledRestoreEDMSampler, sampling.py:697-757): the sliding
  windows are partitioned over the ranks; every step each rank denoises its own windows and
  accumulates their Gaussian-weighted results, then ONLY the parts of those accumulations that
  fall inside another rank's windows (the tile-overlap halos) are exchanged, as point-to-point
  sends between the ranks that actually overlap.  The weight-sum ``count`` is data independent, so
  it is computed locally and never travels.  Every rank draws the same full-latent noise from a
  generator seeded identically (sampling.py:728-731 draws it once per step and slices it per tile).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .sampling import gaussian_weights, sliding_windows

Window = Tuple[int, int, int, int]  # (h0, h1, w0, w1)


def shard_images(num_images: int, rank: int, world: int) -> List[int]:
    """infer_dir-style data parallelism: image i goes to rank i mod world."""
    return list(range(rank, num_images, world))


def partition_windows(windows: Sequence[Window], world: int) -> List[List[Window]]:
    """Contiguous blocks of the row-major window list (neighbouring windows share the most overlap,
    so contiguous ownership minimises the number of peers a rank exchanges halos with).  Ranks
    beyond the number of windows get an empty list."""
    n = len(windows)
    out, start = [], 0
    for r in range(world):
        cnt = n // world + (1 if r < n % world else 0)
        out.append(list(windows[start:start + cnt]))
        start += cnt
    return out


def _intersect(a: Window, b: Window) -> Optional[Window]:
    h0, h1, w0, w1 = max(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), min(a[3], b[3])
    return (h0, h1, w0, w1) if h0 < h1 and w0 < w1 else None


def halo_plan(parts: List[List[Window]]) -> Dict[Tuple[int, int], List[Window]]:
    """plan[(src, dst)] = rectangles of src's accumulation that lie inside dst's windows (deduplicated,
    deterministic order).  Only rank pairs whose windows overlap appear."""
    plan: Dict[Tuple[int, int], List[Window]] = {}
    for s, ws in enumerate(parts):
        for d, wd in enumerate(parts):
            if s == d:
                continue
            rects = []
            for a in ws:
                for b in wd:
                    r = _intersect(a, b)
                    if r is not None and r not in rects:
                        rects.append(r)
            if rects:
                plan[(s, d)] = sorted(rects)
    return plan


class TileShardedStepper:
    """Runs tiled sampler steps with the windows of ONE latent sharded over the process group.

    ``step_fn(x_tile, i, noise_tile, window) -> x_tile_next`` denoises one window (on the GPU this is
    ``Stage2Engine.step`` with the per-tile control latent bound); ``accumulate(tile, weight, acc, h0, w0)``
    adds ``tile * weight`` into ``acc`` (``ops.tile_accumulate`` on the GPU).
    """

    def __init__(self, height: int, width: int, tile: int = 128, stride: int = 96, group=None,
                 device: torch.device = torch.device("cpu")):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.windows = sliding_windows(height, width, tile, stride)
        self.parts = partition_windows(self.windows, self.world)
        self.mine = self.parts[self.rank]
        self.plan = halo_plan(self.parts)
        self.weight = gaussian_weights(tile, tile).to(device)
        # count = sum of weights over ALL windows: data independent, computed locally (sampling.py:754)
        self.count = torch.zeros(1, 1, height, width, device=device)
        for (h0, h1, w0, w1) in self.windows:
            self.count[:, :, h0:h1, w0:w1] += self.weight
        self.halo_bytes_per_step = 0

    def exchange(self, acc: torch.Tensor) -> torch.Tensor:
        """Adds to `acc` the other ranks' contributions inside this rank's windows (halo strips only)."""
        if self.world == 1:
            return acc
        ops_list, recv_bufs = [], []
        nbytes = 0
        for (s, d), rects in self.plan.items():
            for (h0, h1, w0, w1) in rects:
                if s == self.rank:
                    buf = acc[:, :, h0:h1, w0:w1].contiguous()
                    ops_list.append(dist.P2POp(dist.isend, buf, d, self.group))
                    nbytes += buf.numel() * buf.element_size()
                elif d == self.rank:
                    buf = torch.empty_like(acc[:, :, h0:h1, w0:w1]).contiguous()
                    ops_list.append(dist.P2POp(dist.irecv, buf, s, self.group))
                    recv_bufs.append((s, (h0, h1, w0, w1), buf))
        if ops_list:
            for req in dist.batch_isend_irecv(ops_list):
                req.wait()
        self.halo_bytes_per_step = nbytes
        total = acc.clone()
        by_src: Dict[int, torch.Tensor] = {}
        for s, (h0, h1, w0, w1), buf in recv_bufs:
            tmp = by_src.setdefault(s, torch.zeros_like(acc))
            tmp[:, :, h0:h1, w0:w1] = buf  # assignment: rectangles of one source may overlap each other
        for tmp in by_src.values():
            total += tmp
        return total

    def step(self, x: torch.Tensor, i: int, noise: torch.Tensor, step_fn: Callable, accumulate: Callable) -> torch.Tensor:
        """One tiled sampler step.  Returns x_next, valid on this rank's windows (sampling.py:716-756)."""
        acc = torch.zeros_like(x)
        for win in self.mine:
            h0, h1, w0, w1 = win
            xt = step_fn(x[:, :, h0:h1, w0:w1].contiguous(), i, noise[:, :, h0:h1, w0:w1].contiguous(), win)
            accumulate(xt, self.weight, acc, h0, w0)
        total = self.exchange(acc)
        return total / self.count

    def gather_full(self, x_local: torch.Tensor) -> torch.Tensor:
        """Assembles the full latent on every rank from the ranks' own regions (used once, after the last step).
        Each pixel is taken from the lowest rank that owns a window covering it."""
        if self.world == 1:
            return x_local
        mask = torch.zeros_like(self.count)
        for (h0, h1, w0, w1) in self.mine:
            mask[:, :, h0:h1, w0:w1] = 1.0
        stacked = [torch.zeros_like(x_local) for _ in range(self.world)]
        masks = [torch.zeros_like(mask) for _ in range(self.world)]
        dist.all_gather(stacked, (x_local * mask).contiguous(), group=self.group)
        dist.all_gather(masks, mask, group=self.group)
        out = torch.zeros_like(x_local)
        filled = torch.zeros_like(mask)
        for xs, ms in zip(stacked, masks):
            take = ms * (1.0 - filled)
            out += xs * take
            filled = torch.clamp(filled + ms, max=1.0)
        return out
