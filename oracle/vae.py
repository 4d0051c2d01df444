"""ORACLE (test infrastructure) — fp32 restatement of the first stage (SDXL VAE) as pure functions of a state_dict.

Follows, file:line relative to the reference root:
  nonlinearity / Normalize / Upsample / Downsample / ResnetBlock / AttnBlock   sgm/modules/diffusionmodules/model.py:43-199
  Encoder.forward / Decoder.forward                                            model.py:573-598, :712-743
  AutoencoderKL.encode / decode                                                sgm/models/autoencoder.py:302-316
  DiagonalGaussianDistribution                                                 sgm/modules/distributions/distributions.py:24-45
  encode_first_stage_with_denoise / decode_first_stage                         models/SR_model.py:65-85

Pinned against the real reference classes in oracle/make_golden.py (vae()): encoder moments and decoded image agree
to fp32 round-off; the reference outputs are committed as tests/golden/vae_64.pt.  Only tests/, __graft_entry__.smoke()
and bench.py's baseline legs may import this module; the product never does.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
SCALE_FACTOR = 0.13025  # model_configs/juggernautXL.yaml:6


def _count(sd: SD, prefix: str) -> int:
    idx = {int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)}
    return max(idx) + 1 if idx else 0


def _conv(sd: SD, p: str, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=padding)


def _gn(sd: SD, p: str, x):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], 1e-6)   # Normalize: eps 1e-6, model.py:49-52


def swish(x):
    return x * torch.sigmoid(x)


def resnet_block(sd: SD, p: str, x):
    """model.py:127-148 with temb = None."""
    h = _conv(sd, p + "conv1.", swish(_gn(sd, p + "norm1.", x)))
    h = _conv(sd, p + "conv2.", swish(_gn(sd, p + "norm2.", h)))
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(sd, p + "nin_shortcut.", x, padding=0)
    return x + h


def attn_block(sd: SD, p: str, x):
    """model.py:176-199: single head, softmax(q k^T / sqrt(C)) v over all pixels."""
    h = _gn(sd, p + "norm.", x)
    q, k, v = (_conv(sd, p + n + ".", h, padding=0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))
    att = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out.", o, padding=0)


def encoder(sd: SD, p: str, x):
    """Encoder.forward — model.py:573-598."""
    h = _conv(sd, p + "conv_in.", x)
    levels = _count(sd, p + "down.")
    for i in range(levels):
        for j in range(_count(sd, f"{p}down.{i}.block.")):
            h = resnet_block(sd, f"{p}down.{i}.block.{j}.", h)
            if (f"{p}down.{i}.attn.{j}.norm.weight") in sd:
                h = attn_block(sd, f"{p}down.{i}.attn.{j}.", h)
        if (f"{p}down.{i}.downsample.conv.weight") in sd:
            h = _conv(sd, f"{p}down.{i}.downsample.conv.", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)   # model.py:82-85
    h = resnet_block(sd, p + "mid.block_1.", h)
    if (p + "mid.attn_1.norm.weight") in sd:
        h = attn_block(sd, p + "mid.attn_1.", h)
    h = resnet_block(sd, p + "mid.block_2.", h)
    return _conv(sd, p + "conv_out.", swish(_gn(sd, p + "norm_out.", h)))


def decoder(sd: SD, p: str, z):
    """Decoder.forward — model.py:712-743."""
    h = _conv(sd, p + "conv_in.", z)
    h = resnet_block(sd, p + "mid.block_1.", h)
    if (p + "mid.attn_1.norm.weight") in sd:
        h = attn_block(sd, p + "mid.attn_1.", h)
    h = resnet_block(sd, p + "mid.block_2.", h)
    levels = _count(sd, p + "up.")
    for i in reversed(range(levels)):
        for j in range(_count(sd, f"{p}up.{i}.block.")):
            h = resnet_block(sd, f"{p}up.{i}.block.{j}.", h)
            if (f"{p}up.{i}.attn.{j}.norm.weight") in sd:
                h = attn_block(sd, f"{p}up.{i}.attn.{j}.", h)
        if (f"{p}up.{i}.upsample.conv.weight") in sd:
            h = _conv(sd, f"{p}up.{i}.upsample.conv.", F.interpolate(h, scale_factor=2.0, mode="nearest"))   # model.py:63-67
    return _conv(sd, p + "conv_out.", swish(_gn(sd, p + "norm_out.", h)))


def moments(sd: SD, x, enc: str = "encoder."):
    """quant_conv(encoder(x)) — autoencoder.py:302-308 / SR_model.py:66-71 (enc = "denoise_encoder.")."""
    return _conv(sd, "quant_conv.", encoder(sd, enc, x), padding=0)


def posterior(m, noise: Optional[torch.Tensor] = None):
    """DiagonalGaussianDistribution.sample / .mode — distributions.py:24-45."""
    mean, logvar = torch.chunk(m, 2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def encode_with_denoise(sd: SD, x, enc: str = "denoise_encoder."):
    """encode_first_stage_with_denoise(x, use_sample=False) — SR_model.py:65-78."""
    return SCALE_FACTOR * posterior(moments(sd, x, enc))


def decode(sd: SD, z):
    """AutoencoderKL.decode — autoencoder.py:310-316."""
    return decoder(sd, "decoder.", _conv(sd, "post_quant_conv.", z, padding=0))


def decode_first_stage(sd: SD, z):
    """SR_model.py:81-85."""
    return decode(sd, 1.0 / SCALE_FACTOR * z).float()
