"""ORACLE (test infrastructure) — fp32 restatement of the stage-2 sampler stack around the network:
sigma table, discrete denoiser with eps-scaling, linear CFG, RestoreEDMSampler step with the
first-block cache, tile windows / weights.  See oracle/stage2.py for the import policy and the
parity-pinning statement.  Pure torch, device-agnostic.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

SIGMA_MAX = 14.6146  # RestoreEDMSampler.sigma_max / LinearCFG — sampling.py:538, guiders.py:48


def ddpm_alphas_cumprod(num_timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.0120):
    """LegacyDDPMDiscretization.__init__ — discretizer.py:42-57 with make_beta_schedule util.py:19-32."""
    betas = torch.linspace(linear_start**0.5, linear_end**0.5, num_timesteps, dtype=torch.float64) ** 2
    return np.cumprod(1.0 - betas.numpy(), axis=0)


def ddpm_sigmas(n: int, do_append_zero: bool = True, flip: bool = False, device="cpu") -> torch.Tensor:
    """LegacyDDPMDiscretization.get_sigmas + Discretization.__call__ — discretizer.py:18-23, :59-69."""
    ac = ddpm_alphas_cumprod()
    if n < 1000:
        ts = np.linspace(1000 - 1, 0, n, endpoint=False).astype(int)[::-1]
        ac = ac[ts]
    elif n != 1000:
        raise ValueError
    sig = torch.tensor((1 - ac) / ac, dtype=torch.float32, device=device) ** 0.5
    sig = torch.flip(sig, (0,))
    if do_append_zero:
        sig = torch.cat([sig, sig.new_zeros([1])])
    return torch.flip(sig, (0,)) if flip else sig


class Denoiser:
    """DiscreteDenoiserWithControl with EpsScaling — denoiser.py:31-78, denoiser_scaling.py:16-22."""

    def __init__(self, device="cpu"):
        self.sigmas = ddpm_sigmas(1000, do_append_zero=False, flip=True, device=device)  # ascending table

    def sigma_to_idx(self, sigma: torch.Tensor) -> torch.Tensor:
        return (sigma - self.sigmas[:, None]).abs().argmin(dim=0).view(sigma.shape)

    def __call__(self, network: Callable, x, sigma, cond, control_scale, fbcache_mode="none", partial_info=None):
        sigma = self.sigmas[self.sigma_to_idx(sigma)]
        shape = sigma.shape
        s = sigma[(...,) + (None,) * (x.ndim - sigma.ndim)]
        c_in = 1 / (s**2 + 1.0) ** 0.5
        c_noise = self.sigma_to_idx(s.reshape(shape))
        out = network(x * c_in, c_noise, cond, control_scale, fbcache_mode, partial_info)
        if "stage1" in fbcache_mode:
            return out
        return out * (-s) + x


def cfg_scale(sigma: torch.Tensor, scale: float, scale_min: float) -> torch.Tensor:
    """LinearCFG.scale_schedule — guiders.py:48."""
    return (scale - scale_min) * sigma / SIGMA_MAX + scale_min


def cfg_prepare(x, s, c: Dict, uc: Dict):
    """LinearCFG.prepare_inputs (uncond first) — guiders.py:65-74."""
    out = {}
    for k in c:
        if k in ("vector", "crossattn", "concat", "control", "control_vector", "mask_x"):
            out[k] = torch.cat((uc[k], c[k]), 0)
        else:
            out[k] = c[k]
    return torch.cat([x] * 2), torch.cat([s] * 2), out


def cfg_combine(x, sigma, scale: float, scale_min: float):
    """LinearCFG.__call__ + NoDynamicThresholding — guiders.py:59-63, sampling_utils.py:7-9."""
    x_u, x_c = x.chunk(2)
    sv = cfg_scale(sigma, scale, scale_min)
    return x_u + sv.view(-1, 1, 1, 1) * (x_c - x_u)


def rel_l1(prev: torch.Tensor, cur: torch.Tensor) -> float:
    """are_two_tensors_similar — models/modules/DFBCache.py:98-112."""
    return ((prev - cur).abs().mean() / (prev.abs().mean() + 1e-6)).item()


class CacheState:
    """MyCacheContext — DFBCache.py:59-69."""

    def __init__(self):
        self.prev = None
        self.final_decode = None


class RestoreSampler:
    """RestoreEDMSampler (restore_cfg <= 0 path) — sampling.py:527-694."""

    def __init__(self, num_steps=50, s_churn=5.0, s_noise=1.003, scale=4.0, scale_min=7.5, device="cpu"):
        self.num_steps, self.s_churn, self.s_noise = num_steps, s_churn, s_noise
        self.scale, self.scale_min = scale, scale_min
        self.device = device
        self.trace: List[Tuple[str, float]] = []

    def init_loop(self, x):
        """prepare_sampling_loop — sampling.py:44-55."""
        sigmas = ddpm_sigmas(self.num_steps, device=self.device)
        x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
        return x, x.new_ones([x.shape[0]]), sigmas

    def denoise(self, x, denoiser, sigma, c, uc, control_scale, threshold, cache: Optional[CacheState]):
        """RestoreEDMSampler.denoise — sampling.py:548-596."""
        xin, sin, cin = cfg_prepare(x, sigma, c, uc)
        if threshold <= 0:
            den = denoiser(xin, sin, cin, control_scale, "none", None)
            return cfg_combine(den, sigma, self.scale, self.scale_min), threshold
        info = denoiser(xin, sin, cin, control_scale, "input_stage1", None)
        if cache.prev is not None:
            diff = rel_l1(cache.prev, info["h"])
            use, th = diff < threshold, diff
        else:
            use, th = False, threshold
        if use and cache.final_decode is not None:
            self.trace.append(("hit", th))
            return cache.final_decode, threshold
        cache.prev = info["h"].clone()
        den = denoiser(xin, sin, cin, control_scale, "input_stage2", info)
        den = cfg_combine(den, sigma, self.scale, self.scale_min)
        cache.final_decode = den.clone()
        self.trace.append(("miss", th))
        return den, th

    def step(self, x, i, s_in, sigmas, denoiser, c, uc, control_scale=1.0, threshold=0.1, cache=None, eps_noise=None):
        """RestoreEDMSampler.step + sampler_step — sampling.py:659-694, :598-621 (restore_cfg off)."""
        gamma = min(self.s_churn / (len(sigmas) - 1), 2**0.5 - 1)
        sigma, next_sigma = s_in * sigmas[i], s_in * sigmas[i + 1]
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = (torch.randn_like(x) if eps_noise is None else eps_noise) * self.s_noise
            x = x + eps * ((sigma_hat**2 - sigma**2) ** 0.5).view(-1, 1, 1, 1)
        den, threshold = self.denoise(x, denoiser, sigma_hat, c, uc, control_scale, threshold, cache)
        d = (x - den) / sigma_hat.view(-1, 1, 1, 1)
        x = x + d * (next_sigma - sigma_hat).view(-1, 1, 1, 1)
        return x, threshold


def gaussian_weights(tile_w: int, tile_h: int) -> torch.Tensor:
    """gaussian_weights (var 0.01; note the asymmetric midpoints) — sampling.py:830-847. Returns [h, w] fp64."""
    var = 0.01
    mid = (tile_w - 1) / 2
    xs = [math.exp(-(x - mid) * (x - mid) / (tile_w * tile_w) / (2 * var)) / math.sqrt(2 * math.pi * var)
          for x in range(tile_w)]
    mid = tile_h / 2
    ys = [math.exp(-(y - mid) * (y - mid) / (tile_h * tile_h) / (2 * var)) / math.sqrt(2 * math.pi * var)
          for y in range(tile_h)]
    return torch.tensor(np.outer(ys, xs))


def sliding_windows(h: int, w: int, tile: int, stride: int):
    """_sliding_windows — sampling.py:850-863."""
    his = list(range(0, h - tile + 1, stride))
    if (h - tile) % stride != 0:
        his.append(h - tile)
    wis = list(range(0, w - tile + 1, stride))
    if (w - tile) % stride != 0:
        wis.append(w - tile)
    return [(hi, hi + tile, wi, wi + tile) for hi in his for wi in wis]


def tiled_step(sampler: RestoreSampler, x, i, s_in, sigmas, denoiser, c, uc, tile: int, stride: int, eps_noise,
               control_scale=1.0):
    """One step of TiledRestoreEDMSampler.__call__ — sampling.py:716-756 — with the tile semantics
    SURVEY.md section 5 fixes for the bit-rotted reference loop: threshold <= 0 path per tile, element 0 of
    the returned tuple, one full-latent eps_noise sliced per tile, control sliced per tile."""
    weights = gaussian_weights(tile, tile).to(x.dtype).to(x.device)[None, None]
    x_next, count = torch.zeros_like(x), torch.zeros_like(x)
    lq = c["control"]
    for (h0, h1, w0, w1) in sliding_windows(x.shape[2], x.shape[3], tile, stride):
        ct = dict(c, control=lq[:, :, h0:h1, w0:w1])
        uct = dict(uc, control=lq[:, :, h0:h1, w0:w1])
        xt, _ = sampler.step(x[:, :, h0:h1, w0:w1], i, s_in, sigmas, denoiser, ct, uct, control_scale, 0.0, None,
                             eps_noise[:, :, h0:h1, w0:w1])
        x_next[:, :, h0:h1, w0:w1] += xt * weights
        count[:, :, h0:h1, w0:w1] += weights
    return x_next / count
