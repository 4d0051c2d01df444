"""Drop-in conditioner of the restoration pipeline on the sm_100a kernels: the two text
towers and the size embedders that turn a caption into the ``crossattn`` [B, 77, 2048] / ``vector`` [B, 2816]
conditioning of the stage-2 networks.  Runs once per image; reuses the GEMM / LayerNorm / attention kernels (causal
mask, GELU epilogues).

Mirrored reference code (relative to the reference root) and the third-party modules it instantiates:
  GeneralConditionerWithControl.forward / get_unconditional_conditioning   sgm/modules/encoders/modules.py:121-234
  FrozenCLIPEmbedder (layer "hidden", layer_idx 11)                        modules.py:436-499
      -> transformers.CLIPTextModel ("openai/clip-vit-large-patch14": 12 layers, width 768, 12 heads, quick_gelu)
  FrozenOpenCLIPEmbedder2 (ViT-bigG-14, layer "penultimate", pooled)       modules.py:501-612
      -> open_clip text tower (32 layers, width 1280, 20 heads, GELU, nn.MultiheadAttention with packed in_proj)
  ConcatTimestepEmbedderND / Timestep                                      modules.py:1031-1047, sgm util timestep_embedding
  prepare_condition (size / crop vectors)                                  models/SR_model.py:127-156
Parameter names follow those modules (``transformer.text_model.encoder.layers.N.self_attn.q_proj.weight``,
``model.transformer.resblocks.N.attn.in_proj_weight`` ...), so the reference's checkpoints load unchanged.

Tokenisation (CLIPTokenizer / open_clip.tokenize) needs vocabulary files that are not available offline and is
host-side string work: these modules take TOKEN IDS [B, 77] (int64).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .modules import Packed, _F32, _linear

bf16 = torch.bfloat16


def _ln(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    return ops.layer_norm(x, norm.weight, norm.bias, norm.eps)


def _pad_rows(x: torch.Tensor, mult: int = 8):
    """Token matrices are [B * 77, C]: GEMM row counts are free, nothing to pad (kept for clarity)."""
    return x


# ------------------------------------------------------------------------------------------------
# CLIP-L text tower, transformers.CLIPTextModel layout
# ------------------------------------------------------------------------------------------------
class CLIPAttention(nn.Module, Packed):
    def __init__(self, width: int, heads: int):
        super().__init__()
        self.heads = heads
        self.q_proj, self.k_proj, self.v_proj = nn.Linear(width, width), nn.Linear(width, width), nn.Linear(width, width)
        self.out_proj = nn.Linear(width, width)

    def forward(self, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        w = self._pk("qkv.w", (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight),
                     lambda q, k, v: torch.cat([q, k, v], 0).to(bf16).contiguous())
        b = self._pk("qkv.b", (self.q_proj.bias, self.k_proj.bias, self.v_proj.bias),
                     lambda q, k, v: torch.cat([q, k, v], 0).detach().float().contiguous())
        c = x.shape[-1]
        qkv = ops.gemm(x, w, b)
        o = ops.attention(qkv, qkv, qkv, self.heads, q_col=0, k_col=c, v_col=2 * c, scale=0.125, causal=True)
        return _linear(self, "out", self.out_proj, o, residual=residual)


class CLIPMLP(nn.Module, Packed):
    def __init__(self, width: int, inner: int, act: int):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(width, inner), nn.Linear(inner, width), act

    def forward(self, x, residual):
        return _linear(self, "fc2", self.fc2, _linear(self, "fc1", self.fc1, x, act=self.act), residual=residual)


class CLIPEncoderLayer(nn.Module):
    """Pre-LN block: x += attn(LN1(x)); x += mlp(LN2(x)) (residual adds in the GEMM epilogues)."""

    def __init__(self, width: int, heads: int, inner: int, act: int):
        super().__init__()
        self.self_attn = CLIPAttention(width, heads)
        self.layer_norm1 = nn.LayerNorm(width)
        self.mlp = CLIPMLP(width, inner, act)
        self.layer_norm2 = nn.LayerNorm(width)

    def forward(self, x):
        x = self.self_attn(_ln(self.layer_norm1, x), x)
        return self.mlp(_ln(self.layer_norm2, x), x)


class _CLIPEmbeddings(nn.Module):
    def __init__(self, vocab, width, positions):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.position_embedding = nn.Embedding(positions, width)

    def forward(self, ids):
        return ops.embed_tokens(ids.contiguous(), self.token_embedding.weight.detach(), self.position_embedding.weight.detach())


class _CLIPEncoder(nn.Module):
    def __init__(self, layers, width, heads, inner, act):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(width, heads, inner, act) for _ in range(layers)])


class _CLIPTextTransformer(nn.Module):
    def __init__(self, vocab, width, layers, heads, inner, positions, act):
        super().__init__()
        self.embeddings = _CLIPEmbeddings(vocab, width, positions)
        self.encoder = _CLIPEncoder(layers, width, heads, inner, act)
        self.final_layer_norm = nn.LayerNorm(width)


class CLIPTextModel(nn.Module):
    """transformers.CLIPTextModel for inference: ``hidden_states(ids)[i]`` = output after i layers."""

    def __init__(self, vocab=49408, width=768, layers=12, heads=12, inner=3072, positions=77, act=3):
        super().__init__()
        self.text_model = _CLIPTextTransformer(vocab, width, layers, heads, inner, positions, act)

    def hidden_states(self, ids: torch.Tensor, upto: Optional[int] = None) -> List[torch.Tensor]:
        tm = self.text_model
        x = tm.embeddings(ids)
        out = [x]
        for layer in list(tm.encoder.layers)[:upto]:
            x = layer(x)
            out.append(x)
        return out


class FrozenCLIPEmbedder(nn.Module):
    """modules.py:436-499 with layer="hidden": returns hidden_states[layer_idx] ([B, 77, 768]); the last layer(s) beyond
    layer_idx are never evaluated."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, freeze=True, layer="hidden",
                 layer_idx=11, always_return_pooled=False, _layers=12):
        super().__init__()
        assert layer == "hidden" and layer_idx is not None and not always_return_pooled and 0 <= layer_idx <= _layers
        self.transformer = CLIPTextModel(layers=_layers)   # _layers: reduced depth for fast tests only
        self.max_length, self.layer, self.layer_idx = max_length, layer, layer_idx
        self.input_key = "txt"

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        ops.require_cuda(ids, "b200sr.conditioner.FrozenCLIPEmbedder")
        return self.transformer.hidden_states(ids, upto=self.layer_idx)[self.layer_idx]

    encode = forward


# ------------------------------------------------------------------------------------------------
# OpenCLIP ViT-bigG-14 text tower, open_clip layout
# ------------------------------------------------------------------------------------------------
class _MHA(nn.Module, Packed):
    """nn.MultiheadAttention's parameter layout (packed in_proj)."""

    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)

    def forward(self, x, residual):
        w = self._pk("in.w", (self.in_proj_weight,), ops.pack_linear)
        b = self._pk("in.b", (self.in_proj_bias,), _F32)
        c = x.shape[-1]
        qkv = ops.gemm(x, w, b)
        o = ops.attention(qkv, qkv, qkv, self.heads, q_col=0, k_col=c, v_col=2 * c, scale=0.125, causal=True)
        return _linear(self, "out", self.out_proj, o, residual=residual)


class _OpenCLIPMLP(nn.Module, Packed):
    def __init__(self, width, inner):
        super().__init__()
        self.c_fc, self.gelu, self.c_proj = nn.Linear(width, inner), nn.GELU(), nn.Linear(inner, width)

    def forward(self, x, residual):
        return _linear(self, "proj", self.c_proj, _linear(self, "fc", self.c_fc, x, act=2), residual=residual)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, width, heads, inner):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = _MHA(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = _OpenCLIPMLP(width, inner)

    def forward(self, x):
        x = self.attn(_ln(self.ln_1, x), x)
        return self.mlp(_ln(self.ln_2, x), x)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads, inner):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, inner) for _ in range(layers)])


class _OpenCLIPText(nn.Module, Packed):
    def __init__(self, vocab=49408, width=1280, layers=32, heads=20, inner=5120, positions=77):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.positional_embedding = nn.Parameter(torch.empty(positions, width))
        self.transformer = _Transformer(width, layers, heads, inner)
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, width))
        self.logit_scale = nn.Parameter(torch.ones([]))


class FrozenOpenCLIPEmbedder2(nn.Module):
    """modules.py:501-612 with layer="penultimate", always_return_pooled=True, legacy=False: returns
    (penultimate hidden state [B, 77, 1280], pooled [B, 1280] = ln_final(last)[eot] @ text_projection)."""

    def __init__(self, arch="ViT-bigG-14", version="laion2b_s39b_b160k", device="cuda", max_length=77, freeze=True,
                 layer="penultimate", always_return_pooled=True, legacy=False, _layers=32):
        super().__init__()
        assert arch == "ViT-bigG-14" and layer == "penultimate" and always_return_pooled and not legacy
        self.model = _OpenCLIPText(layers=_layers)   # _layers: reduced depth for fast tests only
        self.max_length, self.layer, self.return_pooled, self.legacy = max_length, layer, always_return_pooled, legacy
        self.input_key = "txt"

    def forward(self, ids: torch.Tensor):
        ops.require_cuda(ids, "b200sr.conditioner.FrozenOpenCLIPEmbedder2")
        m = self.model
        x = ops.embed_tokens(ids.contiguous(), m.token_embedding.weight.detach(), m.positional_embedding.detach())
        blocks = list(m.transformer.resblocks)
        for r in blocks[:-1]:
            x = r(x)
        penultimate = x                                                   # before the last block (modules.py:598-600)
        last = _ln(m.ln_final, blocks[-1](x))
        b = ids.shape[0]
        eot = ids.argmax(dim=-1)                                          # eot token = highest id (modules.py:589-593)
        rows = last[torch.arange(b, device=ids.device), eot].contiguous()   # [B, 1280] gather of B rows
        w = m._pk("proj", (m.text_projection,), lambda t: t.detach().t().to(bf16).contiguous())   # x @ P
        pooled = ops.gemm(rows, w, out_fp32=True)
        return penultimate, pooled

    encode = forward


# ------------------------------------------------------------------------------------------------
# size / crop embedders and the conditioner
# ------------------------------------------------------------------------------------------------
class ConcatTimestepEmbedderND(nn.Module):
    """modules.py:1031-1047: each scalar of x [B, d] -> 256-wide (cos | sin) embedding, concatenated: [B, d * outdim]."""

    def __init__(self, outdim, input_key=None):
        super().__init__()
        self.outdim, self.input_key = outdim, input_key

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim == 1:
            x = x[:, None]
        b, dims = x.shape
        emb = ops.sinusoid_embedding(x.reshape(-1).float(), self.outdim, 10000.0, sin_first=False)
        return emb.view(b, dims * self.outdim)


class GeneralConditionerWithControl(nn.Module):
    """modules.py:72-234 for the shipped emb_models list: crossattn = cat(CLIP-L hidden 11, bigG penultimate) along
    channels, vector = cat(bigG pooled, original_size, crop_coords, target_size embeddings); "control" passes through."""

    KEYS = ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple")

    def __init__(self, _clip_layers=12, _clip_layer_idx=11, _bigg_layers=32):
        super().__init__()
        self.embedders = nn.ModuleList([FrozenCLIPEmbedder(layer_idx=_clip_layer_idx, _layers=_clip_layers),
                                        FrozenOpenCLIPEmbedder2(_layers=_bigg_layers)] +
                                       [ConcatTimestepEmbedderND(256, k) for k in self.KEYS])

    @torch.no_grad()
    def forward(self, batch: Dict, force_zero_embeddings: Optional[Sequence[str]] = None) -> Dict:
        """batch["txt"] = (ids for the CLIP-L tokenizer, ids for the open_clip tokenizer), each int64 [B, 77]."""
        force = set(force_zero_embeddings or [])
        ids_l, ids_g = batch["txt"]
        h_l = self.embedders[0](ids_l)
        h_g, pooled = self.embedders[1](ids_g)
        vec = [pooled] + [self.embedders[2 + i](batch[k]).float() for i, k in enumerate(self.KEYS)]
        cross = torch.cat((h_l, h_g), dim=2)
        if "txt" in force:
            cross, vec[0] = torch.zeros_like(cross), torch.zeros_like(vec[0])
        out = {"crossattn": cross.float(), "vector": torch.cat(vec, dim=1)}
        if "control" in batch:
            out["control"] = batch["control"]
        return out

    def get_unconditional_conditioning(self, batch_c: Dict, batch_uc: Optional[Dict] = None,
                                       force_uc_zero_embeddings: Optional[Sequence[str]] = None):
        """modules.py:163-181."""
        c = self(batch_c)
        uc = self(batch_c if batch_uc is None else batch_uc, force_uc_zero_embeddings or [])
        return c, uc


def prepare_condition(conditioner: GeneralConditionerWithControl, z: torch.Tensor, ids_c, ids_uc, size=(1024, 1024)):
    """SR_backbone.prepare_condition (models/SR_model.py:127-143) on token ids: (c, uc) dicts for Stage2Engine."""
    n, dev = z.shape[0], z.device
    batch = {"original_size_as_tuple": torch.tensor(size, device=dev).repeat(n, 1).float(),
             "crop_coords_top_left": torch.tensor([0, 0], device=dev).repeat(n, 1).float(),
             "target_size_as_tuple": torch.tensor(size, device=dev).repeat(n, 1).float(), "control": z}
    batch_uc = dict(batch, txt=ids_uc)
    batch["txt"] = ids_c
    return conditioner.get_unconditional_conditioning(batch, batch_uc)
