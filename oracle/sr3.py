"""ORACLE (test infrastructure) — fp32 restatement of the SR3 stage-1 UNet and its DDPM ancestral
loop as pure functions of a ``state_dict``.  Import policy and parity-pinning statement: see
oracle/stage2.py.  Reference: models/sr3_model/sr3_modules/{unet,diffusion}.py.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from .stage2 import _conv, _count, _has, _linear

SD = Dict[str, torch.Tensor]


def noise_embedding(sd: SD, p: str, noise_level: torch.Tensor, dim: int) -> torch.Tensor:
    """PositionalEncoding (sin | cos) + MLP — unet.py:19-32, :176-181. noise_level: [B, 1]."""
    count = dim // 2
    step = torch.arange(count, dtype=noise_level.dtype, device=noise_level.device) / count
    enc = noise_level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)  # [B, 1, dim]
    h = _linear(sd, p + "noise_level_mlp.1.", enc)
    h = h * torch.sigmoid(h)
    return _linear(sd, p + "noise_level_mlp.3.", h)


def _block(sd: SD, p: str, x: torch.Tensor, groups: int = 32) -> torch.Tensor:
    """Block: GN -> Swish -> (Dropout: identity in eval) -> conv3x3 — unet.py:81-92."""
    h = F.group_norm(x, groups, sd[p + "block.0.weight"], sd[p + "block.0.bias"], 1e-5)
    return _conv(sd, p + "block.3.", h * torch.sigmoid(h))


def resnet_block(sd: SD, p: str, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """ResnetBlock (additive noise affine) — unet.py:95-111, :35-51."""
    h = _block(sd, p + "block1.", x)
    h = h + _linear(sd, p + "noise_func.noise_func.0.", t).view(x.shape[0], -1, 1, 1)
    h = _block(sd, p + "block2.", h)
    res = _conv(sd, p + "res_conv.", x, padding=0) if _has(sd, p + "res_conv.weight") else x
    return h + res


def self_attention(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """SelfAttention, one head of width C, scale 1/sqrt(C) — unet.py:114-143."""
    b, c, h, w = x.shape
    n = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.conv2d(n, sd[p + "qkv.weight"]).view(b, 1, 3 * c, h * w)
    q, k, v = qkv.chunk(3, dim=2)  # [b, 1, c, hw]
    att = torch.softmax(q.transpose(-1, -2) @ k / math.sqrt(c), dim=-1)  # [b, 1, hw_q, hw_k]
    o = (v @ att.transpose(-1, -2)).view(b, c, h, w)
    return _conv(sd, p + "out.", o, padding=0) + x


def _layer(sd: SD, p: str, x, t):
    if _has(sd, p + "res_block.block1.block.0.weight"):  # ResnetBlocWithAttn — unet.py:146-159
        x = resnet_block(sd, p + "res_block.", x, t)
        if _has(sd, p + "attn.qkv.weight"):
            x = self_attention(sd, p + "attn.", x)
        return x, True
    if _has(sd, p + "conv.weight"):
        if sd[p + "conv.weight"].shape[0] == sd[p + "conv.weight"].shape[1] and p.split(".")[-3] == "ups":
            return _conv(sd, p + "conv.", F.interpolate(x, scale_factor=2, mode="nearest")), False  # Upsample :59-66
        return _conv(sd, p + "conv.", x, stride=2), False  # Downsample :69-75
    return _conv(sd, p, x), False  # first conv — unet.py:190-191


def unet(sd: SD, p: str, x: torch.Tensor, noise_level: torch.Tensor, inner_channel: int = 64) -> torch.Tensor:
    """UNet.forward — unet.py:236-261."""
    t = noise_embedding(sd, p, noise_level, inner_channel)
    feats: List[torch.Tensor] = []
    for i in range(_count(sd, p + "downs.")):
        x, _ = _layer(sd, f"{p}downs.{i}.", x, t)
        feats.append(x)
    for i in range(_count(sd, p + "mid.")):
        x, _ = _layer(sd, f"{p}mid.{i}.", x, t)
    for i in range(_count(sd, p + "ups.")):
        q = f"{p}ups.{i}."
        if _has(sd, q + "res_block.block1.block.0.weight"):
            x, _ = _layer(sd, q, torch.cat((x, feats.pop()), dim=1), t)
        else:
            x, _ = _layer(sd, q, x, t)
    return _block(sd, p + "final_conv.", x)


class Schedule:
    """GaussianDiffusion.set_new_noise_schedule (linear) — diffusion.py:93-140."""

    def __init__(self, n_timestep=50, linear_start=1e-6, linear_end=1e-2):
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.n = n_timestep
        self.sqrt_alphas_cumprod_prev = np.sqrt(np.append(1.0, ac))
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.sqrt_recip = f32(np.sqrt(1.0 / ac))
        self.sqrt_recipm1 = f32(np.sqrt(1.0 / ac - 1))
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.log_var = f32(np.log(np.maximum(post_var, 1e-20)))
        self.coef1 = f32(betas * np.sqrt(ac_prev) / (1.0 - ac))
        self.coef2 = f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac))


def p_sample(denoise_fn, sched: Schedule, x, t: int, cond, noise):
    """p_mean_variance + p_sample — diffusion.py:152-176. `noise` replaces torch.randn_like."""
    level = torch.full((x.shape[0], 1), float(np.float32(sched.sqrt_alphas_cumprod_prev[t + 1])), device=x.device)
    eps = denoise_fn(torch.cat([cond, x], dim=1), level)
    x0 = (sched.sqrt_recip[t] * x - sched.sqrt_recipm1[t] * eps).clamp(-1.0, 1.0)
    mean = sched.coef1[t] * x0 + sched.coef2[t] * x
    if t > 0:
        return mean + noise * (0.5 * sched.log_var[t]).exp()
    return mean


def p_sample_loop(denoise_fn, sched: Schedule, cond, noises: List[torch.Tensor]):
    """p_sample_loop (conditional, continous=False) — diffusion.py:178-201.
    noises[0] is the initial image, noises[1 + k] the noise of the k-th step."""
    img = noises[0]
    for k, t in enumerate(reversed(range(sched.n))):
        img = p_sample(denoise_fn, sched, img, t, cond, noises[1 + k])
    return img
