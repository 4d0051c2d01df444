"""ORACLE (test infrastructure, not product code) — fp32 restatement of the reference stage-2
denoiser networks as pure functions of a ``state_dict``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product path (``b200sr``) never does.

Parity status: the reference ships no tests or golden vectors for this path ("parity unpinned"
by the reference itself).  This restatement is pinned instead against the reference's own
modules imported from /root/reference in the build container (``oracle/make_golden.py``; max abs
difference ~1e-5 in fp32), and the resulting input/output vectors are committed under
``tests/golden/``.

The module tree is *inferred from the state_dict keys* (which sub-module exists at which index),
so the functions below are independent of both the reference's and the product's constructors.
Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _has(sd: SD, key: str) -> bool:
    return key in sd


def _count(sd: SD, prefix: str) -> int:
    """number of consecutive integer-indexed children under `prefix` ("a.b." -> a.b.0, a.b.1, ...)."""
    n = 0
    while any(k.startswith(f"{prefix}{n}.") for k in sd):
        n += 1
    return n


def _linear(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def _conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=padding)


def _gn(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    # GroupNorm32 = nn.GroupNorm(32, C) — sgm/modules/diffusionmodules/util.py:258-276
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], eps)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """cos | sin sinusoid — sgm/modules/diffusionmodules/util.py:206-230."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def embed(sd: SD, p: str, t: torch.Tensor, y: torch.Tensor, model_channels: int) -> torch.Tensor:
    """time_embed(t_emb) + label_emb(y) — openaimodel.py:657-691, :987-992; SR_modules.py:512-519."""
    e = timestep_embedding(t, model_channels)
    e = _linear(sd, p + "time_embed.2.", F.silu(_linear(sd, p + "time_embed.0.", e)))
    l = _linear(sd, p + "label_emb.0.2.", F.silu(_linear(sd, p + "label_emb.0.0.", y)))
    return e + l


def resblock(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward (no up/down, no scale-shift) — openaimodel.py:324-350."""
    h = _conv(sd, p + "in_layers.2.", F.silu(_gn(sd, p + "in_layers.0.", x)))
    e = _linear(sd, p + "emb_layers.1.", F.silu(emb))
    h = h + e[:, :, None, None]
    h = _conv(sd, p + "out_layers.3.", F.silu(_gn(sd, p + "out_layers.0.", h)))
    if _has(sd, p + "skip_connection.weight"):
        x = _conv(sd, p + "skip_connection.", x, padding=0)
    return x + h


# "manual" = softmax(q k^T / 8) v spelled out (what the golden vectors were pinned with); "sdpa" = the call the reference
# itself makes without xformers, F.scaled_dot_product_attention (attention.py:275-277) — selected by bench.py's
# GPU eager baseline so that stock torch runs its fused attention kernel instead of materialising the score matrix.
ATTENTION = "manual"


def cross_attention(sd: SD, p: str, x: torch.Tensor, context: Optional[torch.Tensor]) -> torch.Tensor:
    """CrossAttention.forward, heads of 64, scale 1/8 — sgm/modules/attention.py:222-285."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    b, n, c = q.shape
    h = c // 64
    q, k, v = (t.reshape(b, -1, h, 64).transpose(1, 2) for t in (q, k, v))
    if ATTENTION == "sdpa":
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    else:
        att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(b, n, c)
    return _linear(sd, p + "to_out.0.", o)


def transformer_block(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """BasicTransformerBlock._forward — attention.py:465-486; GEGLU / FeedForward :84-110."""

    def ln(name, t):
        return F.layer_norm(t, (t.shape[-1],), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5)

    x = cross_attention(sd, p + "attn1.", ln("norm1", x), None) + x
    x = cross_attention(sd, p + "attn2.", ln("norm2", x), context) + x
    val, gate = _linear(sd, p + "ff.net.0.proj.", ln("norm3", x)).chunk(2, dim=-1)
    x = _linear(sd, p + "ff.net.2.", val * F.gelu(gate)) + x
    return x


def spatial_transformer(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """SpatialTransformer.forward with use_linear=True — attention.py:614-635 (norm eps 1e-6, :122-125)."""
    b, c, h, w = x.shape
    t = _gn(sd, p + "norm.", x, eps=1e-6).permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = _linear(sd, p + "proj_in.", t)
    for d in range(_count(sd, p + "transformer_blocks.")):
        t = transformer_block(sd, f"{p}transformer_blocks.{d}.", t, context)
    t = _linear(sd, p + "proj_out.", t)
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2) + x


def layer(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """One child of a TimestepEmbedSequential — dispatch of openaimodel.py:92-98."""
    if _has(sd, p + "in_layers.0.weight"):
        return resblock(sd, p, x, emb)
    if _has(sd, p + "transformer_blocks.0.norm1.weight"):
        return spatial_transformer(sd, p, x, context)
    if _has(sd, p + "op.weight"):  # Downsample — openaimodel.py:190-204
        return _conv(sd, p + "op.", x, stride=2)
    if _has(sd, p + "conv.weight"):  # Upsample — openaimodel.py:125-145
        return _conv(sd, p + "conv.", F.interpolate(x, scale_factor=2, mode="nearest"))
    if _has(sd, p + "weight"):  # plain 3x3 conv (input_blocks.0.0, input_hint_block.0)
        return _conv(sd, p, x)
    raise KeyError(f"oracle: cannot classify module at {p}")


def sequential(sd: SD, p: str, x, emb, context):
    for j in range(_count(sd, p)):
        x = layer(sd, f"{p}{j}.", x, emb, context)
    return x


def zero_sft(sd: SD, p: str, c: torch.Tensor, h: torch.Tensor, h_ori: Optional[torch.Tensor], control_scale: float):
    """ZeroSFT.forward — models/modules/SR_modules.py:88-110."""
    pre_concat = sd[p + "param_free_norm.weight"].shape[0] != sd[p + "zero_conv.weight"].shape[0]
    h_raw = torch.cat([h_ori, h], dim=1) if (h_ori is not None and pre_concat) else h
    h = h + _conv(sd, p + "zero_conv.", c, padding=0)
    if h_ori is not None and pre_concat:
        h = torch.cat([h_ori, h], dim=1)
    actv = F.silu(_conv(sd, p + "mlp_shared.0.", c))
    gamma = _conv(sd, p + "zero_mul.", actv)
    beta = _conv(sd, p + "zero_add.", actv)
    h = _gn(sd, p + "param_free_norm.", h) * (gamma + 1) + beta
    if h_ori is not None and not pre_concat:
        h = torch.cat([h_ori, h], dim=1)
    return h * control_scale + h_raw * (1 - control_scale)


def zero_cross_attn(sd: SD, p: str, context: torch.Tensor, x: torch.Tensor, control_scale: float):
    """ZeroCrossAttn.forward — SR_modules.py:135-149."""
    b, c, h, w = x.shape
    xt = _gn(sd, p + "norm1.", x).permute(0, 2, 3, 1).reshape(b, h * w, c)
    ct = _gn(sd, p + "norm2.", context).permute(0, 2, 3, 1).reshape(b, h * w, -1)
    o = cross_attention(sd, p + "attn.", xt, ct)
    return x + o.reshape(b, h, w, c).permute(0, 3, 1, 2) * control_scale


def adapter(sd: SD, i: int, p: str, control, h, h_ori=None, control_scale=1.0):
    q = f"{p}project_modules.{i}."
    if _has(sd, q + "zero_conv.weight"):
        return zero_sft(sd, q, control, h, h_ori, control_scale)
    assert h_ori is None
    return zero_cross_attn(sd, q, control, h, control_scale)


def glv_control(sd: SD, p: str, x, timesteps, xt, context, y, model_channels: int = 320) -> List[torch.Tensor]:
    """GLVControl.forward — SR_modules.py:496-537."""
    emb = embed(sd, p, timesteps, y, model_channels)
    hint = sequential(sd, p + "input_hint_block.", x, emb, context)
    hs = []
    h = xt
    for i in range(_count(sd, p + "input_blocks.")):
        h = sequential(sd, f"{p}input_blocks.{i}.", h, emb, context)
        if i == 0:
            h = h + hint
        hs.append(h)
    h = sequential(sd, p + "middle_block.", h, emb, context)
    hs.append(h)
    return hs


def unet_input_stage(sd: SD, p: str, x, timesteps, context, y, model_channels: int = 320):
    """LightGLVUNet.forward, shared prologue + input blocks — SR_modules.py:611-627 / :660-684."""
    emb = embed(sd, p, timesteps, y, model_channels)
    hs = []
    h = x
    for i in range(_count(sd, p + "input_blocks.")):
        h = sequential(sd, f"{p}input_blocks.{i}.", h, emb, context)
        hs.append(h)
    return h, hs, emb


def unet_output_stage(sd: SD, p: str, h, hs, emb, context, control, control_scale: float):
    """middle block, adapters and output blocks — SR_modules.py:628-657 / :699-730."""
    hs = list(hs)
    a = _count(sd, p + "project_modules.") - 1
    ci = len(control) - 1
    h = sequential(sd, p + "middle_block.", h, emb, context)
    h = adapter(sd, a, p, control[ci], h, None, control_scale)
    a -= 1
    ci -= 1
    for i in range(_count(sd, p + "output_blocks.")):
        q = f"{p}output_blocks.{i}."
        h = adapter(sd, a, p, control[ci], hs.pop(), h, control_scale)
        a -= 1
        if _count(sd, q) == 3:
            h = layer(sd, q + "0.", h, emb, context)
            h = layer(sd, q + "1.", h, emb, context)
            h = adapter(sd, a, p, control[ci], h, None, control_scale)
            a -= 1
            h = layer(sd, q + "2.", h, emb, context)
        else:
            h = sequential(sd, q, h, emb, context)
        ci -= 1
    # self.out = GN32 + SiLU + conv3x3 — openaimodel.py:941-947
    return _conv(sd, p + "out.2.", F.silu(_gn(sd, p + "out.0.", h)))


def control_wrapper(sd: SD, x, t, c: dict, control_scale: float = 1.0, model_channels: int = 320):
    """ControlWrapper.forward, fbcache_mode="none" — sgm/modules/diffusionmodules/wrappers.py:84-110
    (fp32: autocast is a no-op on the oracle)."""
    control = glv_control(sd, "control_model.", c["control"], t, x, c["crossattn"], c["vector"], model_channels)
    h, hs, emb = unet_input_stage(sd, "diffusion_model.", x, t, c["crossattn"], c["vector"], model_channels)
    return unet_output_stage(sd, "diffusion_model.", h, hs, emb, c["crossattn"], control, control_scale).float()


def network(sd: SD):
    """The callable DiscreteDenoiserWithControl hands its arguments to (denoiser.py:76-78): ControlWrapper.forward
    for fbcache_mode "none" / "input_stage1" / "input_stage2" (wrappers.py:84-110, SR_modules.py:660-730).  The
    control features computed in stage 1 are carried in partial_info, as the reference does (SR_modules.py:694)."""

    def net(x, t, c, control_scale=1.0, fbcache_mode="none", partial_info=None):
        with torch.no_grad():
            if fbcache_mode == "none":
                return control_wrapper(sd, x, t, c, control_scale)
            if fbcache_mode == "input_stage1":
                control = glv_control(sd, "control_model.", c["control"], t, x, c["crossattn"], c["vector"])
                h, hs, emb = unet_input_stage(sd, "diffusion_model.", x, t, c["crossattn"], c["vector"])
                return {"mode": "input", "h": h, "hs": hs, "emb": emb, "context": c["crossattn"], "control": control}
            if fbcache_mode == "input_stage2":
                p = partial_info
                return unet_output_stage(sd, "diffusion_model.", p["h"], p["hs"], p["emb"], p["context"], p["control"],
                                         control_scale).float()
        raise ValueError(fbcache_mode)

    return net
