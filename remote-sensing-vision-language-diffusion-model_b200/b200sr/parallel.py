"""Multi-GPU sharding of the denoiser path: one process per GPU (torch.distributed; NCCL over
NVLink on the B200 box, gloo in the CPU tests).

The path shards two ways and never needs model parallelism (7.7 GB of bf16
weights fit on every GPU):

* images (infer_dir.py:198-202 processes them in a sequential loop): ``shard_images`` deals them
  round-robin, no communication during sampling;
* latent tiles (TiledRestoreEDMSampler, sampling.py:697-757): the sliding windows of ALL images of a
  pool form one work list of (image, window) units, cut into contiguous blocks per rank (an image
  whose windows straddle a cut is shared by two ranks).  Every step each rank denoises its units;
  then ONLY the parts of a window's Gaussian-weighted result that fall inside a window owned by
  another rank (the tile-overlap halos) travel, as one point-to-point message per overlapping rank
  pair.  The weight sum ``count`` is data independent, so it is computed locally and never travels.
  Every rank draws the same full-latent noise per (image, step) from an identically seeded generator
  (sampling.py:728-731 draws it once per step and slices it per tile).

Bit-exactness: the reference accumulates ``x_next[win] += tile * w`` window by window.  Here every
rank adds the contributions to its region in that same global window order — its own tiles through
``tile_accumulate``, foreign ones as received strips through ``strip_add`` — and both sides round the
product and the sum separately, so the N-rank result equals the 1-rank result bit for bit.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .sampling import gaussian_weights, sliding_windows

Window = Tuple[int, int, int, int]  # (h0, h1, w0, w1)
Unit = Tuple[int, int]              # (image index, window index)


def shard_images(num_images: int, rank: int, world: int) -> List[int]:
    """infer_dir-style data parallelism: image i goes to rank i mod world."""
    return list(range(rank, num_images, world))


def partition_windows(windows: Sequence, world: int) -> List[list]:
    """Contiguous blocks of a row-major work list (neighbouring windows share the most overlap, so contiguous
    ownership minimises the number of peers a rank exchanges halos with).  Ranks beyond the number of items
    get an empty list."""
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


def disjoint_cover(rects: Sequence[Window]) -> List[Window]:
    """Non-overlapping rectangles whose union equals the union of `rects` (deterministic order).  The pieces of
    one window that several windows of a peer need overlap each other; each pixel must travel and be added once."""
    rects = [r for r in rects if r is not None]
    if not rects:
        return []
    hs = sorted({r[0] for r in rects} | {r[1] for r in rects})
    ws = sorted({r[2] for r in rects} | {r[3] for r in rects})
    bands = []  # (h0, h1, [column spans])
    for h0, h1 in zip(hs[:-1], hs[1:]):
        spans = []
        for w0, w1 in zip(ws[:-1], ws[1:]):
            if any(r[0] <= h0 and h1 <= r[1] and r[2] <= w0 and w1 <= r[3] for r in rects):
                if spans and spans[-1][1] == w0:
                    spans[-1] = (spans[-1][0], w1)
                else:
                    spans.append((w0, w1))
        if not spans:
            continue
        if bands and bands[-1][1] == h0 and bands[-1][2] == spans:
            bands[-1] = (bands[-1][0], h1, spans)
        else:
            bands.append((h0, h1, spans))
    return [(h0, h1, w0, w1) for h0, h1, spans in bands for (w0, w1) in spans]


class _TorchBlend:
    """Plain-torch blend primitives (CPU tests; the GPU path passes ``b200sr.ops``)."""

    @staticmethod
    def tile_accumulate(tile, weight, acc, cnt, h0, w0):
        th, tw = tile.shape[-2:]
        acc[:, :, h0:h0 + th, w0:w0 + tw] += tile * weight

    @staticmethod
    def tile_weighted_strip(tile, weight, y0, x0, sh, sw, out=None):
        r = (tile * weight)[:, :, y0:y0 + sh, x0:x0 + sw]
        if out is not None:
            out.copy_(r)
            return out
        return r.contiguous()

    @staticmethod
    def strip_add(strip, acc, h0, w0):
        sh, sw = strip.shape[-2:]
        acc[:, :, h0:h0 + sh, w0:w0 + sw] += strip

    @staticmethod
    def tile_normalize(acc, cnt):
        return acc / cnt


class PooledTileStepper:
    """Tiled sampler steps for a POOL of images with the (image, window) work list sharded over the ranks.

    ``step_fn(img, wins, x_tiles, i, noise_tiles) -> x_tiles_next`` denoises a group of windows of one image
    (``x_tiles``: the windows' latents stacked along the batch; on the GPU this is ``Stage2Engine.step`` with the
    windows' control slices bound).  ``blend`` provides tile_accumulate / tile_weighted_strip / strip_add /
    tile_normalize (``b200sr.ops`` on the GPU).
    """

    def __init__(self, n_images: int, height: int, width: int, tile: int = 128, stride: int = 96, group=None,
                 device: torch.device = torch.device("cpu"), tile_batch: int = 1, channels: int = 4, batch: int = 1,
                 blend=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_images, self.H, self.W, self.tile = n_images, height, width, tile
        self.device = torch.device(device)
        self.tile_batch = max(1, tile_batch)
        self.blend = blend if blend is not None else _TorchBlend
        self.windows = sliding_windows(height, width, tile, stride)
        nw = len(self.windows)
        self.units: List[Unit] = [(m, k) for m in range(n_images) for k in range(nw)]
        self.parts: List[List[Unit]] = partition_windows(self.units, self.world)
        self.mine: List[Unit] = self.parts[self.rank]
        self.owner = {u: r for r, part in enumerate(self.parts) for u in part}
        self.my_images = sorted({m for m, _ in self.mine})
        self.weight = gaussian_weights(tile, tile).to(self.device)
        # count = sum of weights over ALL windows: data independent, computed locally (sampling.py:754)
        self.count = torch.zeros(1, 1, height, width, device=self.device)
        for (h0, h1, w0, w1) in self.windows:
            self.count[:, :, h0:h1, w0:w1] += self.weight
        # halo plan: strips[(src, dst)] = [(image, window index of src, rectangle)] in (image, window, rectangle)
        # order — the order both sides pack / unpack the pair's message in
        self.strips: Dict[Tuple[int, int], List[Tuple[int, int, Window]]] = {}
        for (m, k) in self.units:
            s = self.owner[(m, k)]
            for d in range(self.world):
                if d == s:
                    continue
                need = [_intersect(self.windows[k], self.windows[k2]) for (m2, k2) in self.parts[d] if m2 == m]
                for rect in disjoint_cover(need):
                    self.strips.setdefault((s, d), []).append((m, k, rect))
        self._elems = channels * batch  # fp32 elements per pixel of a latent (batch x channels)
        self.halo_bytes_per_step = 4 * self._elems * sum(
            (r[1] - r[0]) * (r[3] - r[2]) for (s, d), lst in self.strips.items() if s == self.rank for _, _, r in lst)
        self.exchange_seconds = 0.0   # accumulated wall time of the send/recv section (diagnostic)
        self._bufs: Dict[Tuple[str, int], torch.Tensor] = {}

    # -- helpers ------------------------------------------------------------------------------------
    def _flat(self, kind: str, peer: int, lst, nb: int, c: int) -> torch.Tensor:
        n = nb * c * sum((r[1] - r[0]) * (r[3] - r[2]) for _, _, r in lst)
        buf = self._bufs.get((kind, peer))
        if buf is None or buf.numel() != n:
            buf = torch.empty(n, dtype=torch.float32, device=self.device)
            self._bufs[(kind, peer)] = buf
        return buf

    @staticmethod
    def _views(buf: torch.Tensor, lst, nb: int, c: int):
        out, off = [], 0
        for (m, k, r) in lst:
            sh, sw = r[1] - r[0], r[3] - r[2]
            n = nb * c * sh * sw
            out.append(((m, k, r), buf[off:off + n].view(nb, c, sh, sw)))
            off += n
        return out

    # -- one tiled sampler step for every image this rank touches ---------------------------------------
    def step(self, xs: Dict[int, torch.Tensor], i: int, noises: Dict[int, torch.Tensor], step_fn: Callable) -> Dict[int, torch.Tensor]:
        """xs / noises: {image: [nb, C, H, W] fp32} for the images of ``my_images``.  Returns the next latents,
        valid on this rank's windows (sampling.py:716-756)."""
        B = self.blend
        tiles: Dict[Unit, torch.Tensor] = {}
        for m in self.my_images:
            wins = [k for (mm, k) in self.mine if mm == m]
            x, noise = xs[m], noises[m]
            nb = x.shape[0]
            for g0 in range(0, len(wins), self.tile_batch):
                grp = wins[g0:g0 + self.tile_batch]
                cut = lambda t: torch.cat([t[:, :, self.windows[k][0]:self.windows[k][1],   # noqa: E731
                                             self.windows[k][2]:self.windows[k][3]] for k in grp], 0).contiguous()
                out = step_fn(m, [self.windows[k] for k in grp], cut(x), i, cut(noise))
                for j, k in enumerate(grp):
                    tiles[(m, k)] = out[j * nb:(j + 1) * nb]
        any_x = next(iter(xs.values())) if xs else None
        nb, c = (any_x.shape[0], any_x.shape[1]) if any_x is not None else (1, 4)
        # ---- halo exchange: one message per overlapping rank pair -------------------------------------
        recv_views: Dict[Tuple[int, int], list] = {}
        if self.world > 1:
            p2p = []
            for (s, d), lst in sorted(self.strips.items()):
                if s == self.rank:
                    buf = self._flat("send", d, lst, nb, c)
                    for (m, k, r), view in self._views(buf, lst, nb, c):
                        h0, _, w0, _ = self.windows[k]
                        B.tile_weighted_strip(tiles[(m, k)], self.weight, r[0] - h0, r[2] - w0, r[1] - r[0], r[3] - r[2],
                                              out=view)
                    p2p.append(dist.P2POp(dist.isend, buf, d, self.group))
                elif d == self.rank:
                    buf = self._flat("recv", s, lst, nb, c)
                    for key, view in self._views(buf, lst, nb, c):
                        recv_views.setdefault((key[0], key[1]), []).append((key[2], view))
                    p2p.append(dist.P2POp(dist.irecv, buf, s, self.group))
            if p2p:
                for req in dist.batch_isend_irecv(p2p):
                    req.wait()
        # ---- blend in global window order ----------------------------------------------------------------
        out: Dict[int, torch.Tensor] = {}
        for m in self.my_images:
            acc = torch.zeros_like(xs[m])
            for k, (h0, h1, w0, w1) in enumerate(self.windows):
                if (m, k) in tiles:
                    B.tile_accumulate(tiles[(m, k)], self.weight, acc, None, h0, w0)
                else:
                    for (r, view) in recv_views.get((m, k), []):
                        B.strip_add(view, acc, r[0], r[2])
            out[m] = B.tile_normalize(acc, self._count_like(acc))
        return out

    def _count_like(self, acc: torch.Tensor) -> torch.Tensor:
        full = self._bufs.get(("count", acc.shape))
        if full is None:
            full = self._bufs[("count", acc.shape)] = self.count.expand_as(acc).contiguous()
        return full

    def region_mask(self, m: int) -> torch.Tensor:
        mask = torch.zeros_like(self.count)
        for (mm, k) in self.mine:
            if mm == m:
                h0, h1, w0, w1 = self.windows[k]
                mask[:, :, h0:h1, w0:w1] = 1.0
        return mask

    def gather_image(self, m: int, x_local: Optional[torch.Tensor], like: torch.Tensor) -> torch.Tensor:
        """Assembles image m's full latent on every rank from the owners' regions (used once, after the last
        step).  Each pixel is taken from the lowest rank that owns a window covering it (all owners hold the same
        bits).  Ranks that do not touch the image pass x_local=None."""
        if self.world == 1:
            return x_local
        mask = self.region_mask(m)
        xl = torch.zeros_like(like) if x_local is None else x_local * mask
        stacked = [torch.zeros_like(like) for _ in range(self.world)]
        masks = [torch.zeros_like(mask) for _ in range(self.world)]
        dist.all_gather(stacked, xl.contiguous(), group=self.group)
        dist.all_gather(masks, mask, group=self.group)
        out = torch.zeros_like(like)
        filled = torch.zeros_like(mask)
        for xs_, ms in zip(stacked, masks):
            take = ms * (1.0 - filled)
            out = torch.where(take.expand_as(out) > 0, xs_, out)
            filled = torch.clamp(filled + ms, max=1.0)
        return out


class TileShardedStepper(PooledTileStepper):
    """The windows of ONE latent sharded over the process group (config 4, single-image scaling).

    ``step_fn(x_tile, i, noise_tile, window) -> x_tile_next`` denoises one window."""

    def __init__(self, height: int, width: int, tile: int = 128, stride: int = 96, group=None,
                 device: torch.device = torch.device("cpu"), blend=None):
        super().__init__(1, height, width, tile, stride, group, device, tile_batch=1, blend=blend)
        self.plan = sorted({(s, d) for (s, d) in self.strips})

    def step(self, x: torch.Tensor, i: int, noise: torch.Tensor, step_fn: Callable, accumulate=None) -> torch.Tensor:
        if not self.mine:
            return x
        fn = lambda m, wins, xt, i_, nt: step_fn(xt, i_, nt, wins[0])  # noqa: E731
        return super().step({0: x}, i, {0: noise}, fn)[0]

    def gather_full(self, x_local: torch.Tensor) -> torch.Tensor:
        return self.gather_image(0, x_local if self.mine else None, x_local)


class EngineTileRunner:
    """``step_fn`` of PooledTileStepper on top of ``Stage2Engine``: per image a caption (c, uc without
    "control") and an LQ latent; the windows' control slices are converted once; captions are switched through
    the engine's snapshots, so a rank that owns windows of several images pays the text folding once per image,
    not once per step."""

    def __init__(self, make_engine: Callable[[], object], captions: Dict[int, Tuple[dict, dict]], lqs: Dict[int, torch.Tensor]):
        self.make_engine = make_engine
        self.captions, self.lqs = captions, lqs
        self.engines: Dict[int, object] = {}      # batch size (windows per call) -> engine
        self.controls: Dict[tuple, torch.Tensor] = {}
        self.current: Dict[int, int] = {}         # engine batch -> image whose caption is live

    def close(self) -> None:
        for eng in self.engines.values():
            eng.close()
        self.engines.clear()
        self.controls.clear()
        self.current.clear()

    def __call__(self, m: int, wins: List[Window], x_tiles: torch.Tensor, i: int, noise_tiles: torch.Tensor) -> torch.Tensor:
        nb = len(wins)
        eng = self.engines.get(nb)
        if eng is None:
            eng = self.engines[nb] = self.make_engine()
        key = (m, tuple(wins))
        ctl = self.controls.get(key)
        lq = self.lqs[m]
        if ctl is None:
            lqt = torch.cat([lq[:, :, h0:h1, w0:w1] for (h0, h1, w0, w1) in wins], 0).contiguous()
            ctl = self.controls[key] = eng.prepare_control(lqt)
        if self.current.get(nb) != m:
            c, uc = self.captions[m]
            lqt = torch.cat([lq[:, :, h0:h1, w0:w1] for (h0, h1, w0, w1) in wins], 0).contiguous()
            eng.set_condition(dict(c, control=lqt), dict(uc, control=lqt), key=m)
            self.current[nb] = m
        out, _ = eng.step(x_tiles, i, noise_tiles, 0.0, control=ctl, copy_out=True)
        return out
