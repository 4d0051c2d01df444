"""ORACLE (test infrastructure) — fp32 restatement of the conditioner as pure functions
of a state_dict.  The text-tower arithmetic lives in third-party packages the reference instantiates
(sgm/modules/encoders/modules.py:436-612): transformers' CLIPTextModel ("openai/clip-vit-large-patch14") and
open_clip's ViT-bigG-14 text tower (open_clip_torch, absent from this image: its published algorithm is restated —
token + positional embedding, pre-LN residual attention blocks with nn.MultiheadAttention (packed in_proj), causal mask,
GELU MLP, ln_final, pooling at the highest token id, text_projection).

Pinned in oracle/make_golden.py (conditioner()) against the installed transformers implementation: CLIPTextModel for the
CLIP-L tower directly, and CLIPTextModelWithProjection configured as the bigG text tower (weights re-laid-out from the
open_clip names) as an independent implementation of the same function.  Golden: tests/golden/conditioner_small.pt.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _count(sd: SD, prefix: str) -> int:
    idx = {int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)}
    return max(idx) + 1 if idx else 0


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def _causal_attention(q, k, v, heads):
    b, t, c = q.shape
    q, k, v = (z.reshape(b, t, heads, c // heads).transpose(1, 2) for z in (q, k, v))
    s = q @ k.transpose(-1, -2) / math.sqrt(c // heads)
    s = s + torch.full((t, t), float("-inf"), device=q.device).triu(1)
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(b, t, c)


def clip_l_hidden_states(sd: SD, p: str, ids: torch.Tensor, heads: int = 12):
    """transformers CLIPTextTransformer: embeddings, then pre-LN layers with quick_gelu; returns hidden_states list."""
    x = sd[p + "embeddings.token_embedding.weight"][ids] + sd[p + "embeddings.position_embedding.weight"][None, : ids.shape[1]]
    out = [x]
    for i in range(_count(sd, p + "encoder.layers.")):
        q = f"{p}encoder.layers.{i}."
        h = _ln(sd, q + "layer_norm1.", x)
        lin = lambda n, t: F.linear(t, sd[q + n + ".weight"], sd[q + n + ".bias"])  # noqa: E731
        a = _causal_attention(lin("self_attn.q_proj", h), lin("self_attn.k_proj", h), lin("self_attn.v_proj", h), heads)
        x = x + lin("self_attn.out_proj", a)
        h = lin("mlp.fc1", _ln(sd, q + "layer_norm2.", x))
        x = x + lin("mlp.fc2", h * torch.sigmoid(1.702 * h))
        out.append(x)
    return out


def openclip_text(sd: SD, p: str, ids: torch.Tensor, heads: int = 20):
    """FrozenOpenCLIPEmbedder2.encode_with_transformer (modules.py:569-612), legacy=False: (penultimate, pooled)."""
    x = sd[p + "token_embedding.weight"][ids] + sd[p + "positional_embedding"][None, : ids.shape[1]]
    n = _count(sd, p + "transformer.resblocks.")
    pen = None
    for i in range(n):
        if i == n - 1:
            pen = x
        q = f"{p}transformer.resblocks.{i}."
        h = _ln(sd, q + "ln_1.", x)
        qkv = F.linear(h, sd[q + "attn.in_proj_weight"], sd[q + "attn.in_proj_bias"])
        a = _causal_attention(*qkv.chunk(3, dim=-1), heads)
        x = x + F.linear(a, sd[q + "attn.out_proj.weight"], sd[q + "attn.out_proj.bias"])
        h = F.linear(_ln(sd, q + "ln_2.", x), sd[q + "mlp.c_fc.weight"], sd[q + "mlp.c_fc.bias"])
        x = x + F.linear(F.gelu(h), sd[q + "mlp.c_proj.weight"], sd[q + "mlp.c_proj.bias"])
    last = _ln(sd, p + "ln_final.", x)
    pooled = last[torch.arange(ids.shape[0], device=ids.device), ids.argmax(dim=-1)] @ sd[p + "text_projection"]
    return pen, pooled


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0):
    """sgm util timestep_embedding (cos | sin) — sgm/modules/diffusionmodules/util.py:206-230."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def concat_timestep(x: torch.Tensor, outdim: int = 256):
    """ConcatTimestepEmbedderND.forward — modules.py:1039-1047."""
    b, d = x.shape
    return timestep_embedding(x.reshape(-1), outdim).reshape(b, d * outdim)


def conditioner(sd: SD, batch: dict, clip_layer_idx: int = 11, zero_txt: bool = False):
    """GeneralConditionerWithControl.forward for the shipped emb_models (modules.py:184-234)."""
    ids_l, ids_g = batch["txt"]
    h_l = clip_l_hidden_states(sd, "embedders.0.transformer.text_model.", ids_l)[clip_layer_idx]
    h_g, pooled = openclip_text(sd, "embedders.1.model.", ids_g)
    cross = torch.cat((h_l, h_g), dim=2)
    if zero_txt:
        cross, pooled = torch.zeros_like(cross), torch.zeros_like(pooled)
    vec = torch.cat([pooled] + [concat_timestep(batch[k].float()) for k in
                                ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple")], dim=1)
    return {"crossattn": cross, "vector": vec}
