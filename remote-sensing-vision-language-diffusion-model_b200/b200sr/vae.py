"""Drop-in first stage (SDXL VAE) of the restoration pipeline on the sm_100a kernels: the encoder /
decoder either side of the stage-2 sampler loop.

Mirrored reference code (relative to the reference root), same constructor arguments, ``forward``
signatures and ``state_dict`` keys:
  Normalize / nonlinearity / Upsample / Downsample / ResnetBlock / AttnBlock / make_attn
                                              sgm/modules/diffusionmodules/model.py:43-313
  Encoder / Decoder                           model.py:482-743
  AutoencoderKL / AutoencoderKLInferenceWrapper   sgm/models/autoencoder.py:282-321
  DiagonalGaussianDistribution                sgm/modules/distributions/distributions.py
  encode_first_stage / encode_first_stage_with_denoise / decode_first_stage   models/SR_model.py:57-85
  the denoise_encoder copy                    models/SR_model.py:22

Numerics = the reference's ``ae_dtype = bf16`` autocast policy (infer.py:63): bf16 operands, fp32
accumulation, fp32 GroupNorm statistics and softmax, bf16 activation storage.

Tiling (utils/tilevae.py VAEHook): with ``fast_encoder = fast_decoder = False`` (SR_model.py:99-125) the
reference's tiled pass aggregates GroupNorm statistics over all tiles and therefore computes the same
function as the untiled network; it exists to fit 24 GB cards.  On a B200 the untiled 2048^2 decode needs
< 10 GB of activations, so the first stage runs whole images; the single-head mid attention is evaluated in
query chunks so that its score matrix never exceeds ``ops.SCORE_CHUNK_BYTES``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .modules import Packed, _F32, _conv3x3, _gn, _gn_silu_conv3x3, _linear, from_nhwc, to_nhwc

bf16 = torch.bfloat16


def Normalize(in_channels, num_groups=32):
    """model.py:49-52."""
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _conv_in_small(holder: Packed, name: str, conv: nn.Conv2d, x_nhwc: torch.Tensor) -> torch.Tensor:
    """3x3 conv from <= 8 channels (image / latent -> features): zero-padded to one 64-channel K chunk so it runs
    on the tcgen05 implicit-GEMM path."""
    w = holder._pk(name + ".w64", (conv.weight,), lambda t: ops.pack_conv3x3_padded(t, 64))
    b = holder._pk(name + ".b", (conv.bias,), _F32)
    return ops.conv3x3(ops.pad_channels(x_nhwc, 64), w, b)


class Upsample(nn.Module, Packed):
    """model.py:53-68: nearest x2 (+ 3x3 conv)."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward_nhwc(self, x):
        x = ops.upsample2x(x)
        return _conv3x3(self, "conv", self.conv, x) if self.with_conv else x

    def forward(self, x):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class Downsample(nn.Module, Packed):
    """model.py:70-88: F.pad(x, (0, 1, 0, 1)) then 3x3 conv, stride 2, padding 0."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        assert with_conv, "the first stage uses the learned downsampling convolution"
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward_nhwc(self, x):
        return _conv3x3(self, "conv", self.conv, x, stride=2, pad_lo=0)

    def forward(self, x):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class ResnetBlock(nn.Module, Packed):
    """model.py:90-148 with temb_channels = 0 (the autoencoder passes temb = None)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        assert temb_channels == 0 and not conv_shortcut
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward_nhwc(self, x, temb=None):
        assert temb is None
        h = _gn_silu_conv3x3(self, "conv1", self.norm1, self.conv1, x)
        skip = x if self.in_channels == self.out_channels else _linear(self, "nin", self.nin_shortcut, x)
        return _gn_silu_conv3x3(self, "conv2", self.norm2, self.conv2, h, residual=skip)   # x + h in the epilogue

    def forward(self, x, temb=None):
        return from_nhwc(self.forward_nhwc(to_nhwc(x), temb))


class AttnBlock(nn.Module, Packed):
    """model.py:158-199 (and its xformers twin :202-263): one head of width C, scale C^-0.5, softmax over all H*W keys;
    q / k / v / proj_out are 1x1 convolutions WITH bias."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def forward_nhwc(self, x, **kwargs):
        b, h, w, c = x.shape
        t = h * w
        n = _gn(self.norm, x).view(b, t, c)
        wv = self._pk("v.w", (self.v.weight,), ops.pack_linear)
        bv = self._pk("v.b", (self.v.bias,), _F32)
        outs = []
        for i in range(b):
            tok = n[i]
            q, k = _linear(self, "q", self.q, tok), _linear(self, "k", self.k, tok)
            v_t = ops.gemm(wv, tok, w_dynamic=True)          # [C, T] = (W_v x)^T; the bias b_v is added after P V
            outs.append(ops.single_head_attention(q, k, v_t, c ** -0.5, out_bias=bv))   # rows of P sum to 1: P (V + 1 b^T) = P V + b^T
        o = outs[0].unsqueeze(0) if b == 1 else torch.stack(outs, 0)
        return _linear(self, "proj_out", self.proj_out, o, residual=x.view(b, t, c)).view(b, h, w, c)

    def forward(self, x, **kwargs):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


MemoryEfficientAttnBlock = AttnBlock


def make_attn(in_channels, attn_type="vanilla", attn_kwargs=None):
    """model.py:276-313: both attention flavours the shipped YAML can select compute the same function."""
    if attn_type in ("vanilla", "vanilla-xformers"):
        assert attn_kwargs is None
        return AttnBlock(in_channels)
    if attn_type == "none":
        return nn.Identity(in_channels)
    raise NotImplementedError(f"attn_type {attn_type} is not used by the shipped first-stage config")


class Encoder(nn.Module, Packed):
    """model.py:482-598."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        assert not use_linear_attn
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)

    def forward_nhwc(self, x):
        h = _conv_in_small(self, "conv_in", self.conv_in, x)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block].forward_nhwc(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block].forward_nhwc(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample.forward_nhwc(h)
        h = self.mid.block_1.forward_nhwc(h)
        h = self.mid.attn_1.forward_nhwc(h) if not isinstance(self.mid.attn_1, nn.Identity) else h
        h = self.mid.block_2.forward_nhwc(h)
        return _gn_silu_conv3x3(self, "conv_out", self.norm_out, self.conv_out, h)   # [B, h, w, 2 z] bf16

    def forward(self, x):
        ops.require_cuda(x, "b200sr.vae.Encoder")
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class Decoder(nn.Module, Packed):
    """model.py:600-743."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        assert not use_linear_attn and not give_pre_end and not tanh_out
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.give_pre_end, self.tanh_out = give_pre_end, tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def get_last_layer(self, **kwargs):
        return self.conv_out.weight

    def forward_nhwc(self, z):
        """z: [B, h, w, z_channels] bf16 -> image fp32 NCHW [B, out_ch, H, W]."""
        h = _conv_in_small(self, "conv_in", self.conv_in, z)
        h = self.mid.block_1.forward_nhwc(h)
        h = self.mid.attn_1.forward_nhwc(h) if not isinstance(self.mid.attn_1, nn.Identity) else h
        h = self.mid.block_2.forward_nhwc(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block].forward_nhwc(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block].forward_nhwc(h)
            if i_level != 0:
                h = self.up[i_level].upsample.forward_nhwc(h)
        h = _gn(self.norm_out, h, silu=True)
        w8, b8 = self._pk("conv_out.wb8", (self.conv_out.weight, self.conv_out.bias), ops.pack_conv3x3_few_out)
        return ops.conv3x3_to_nchw_f32(h, w8, b8, self.conv_out.weight.shape[0])

    def forward(self, z, **kwargs):
        ops.require_cuda(z, "b200sr.vae.Decoder")
        self.last_z_shape = z.shape
        return self.forward_nhwc(to_nhwc(z))


class DiagonalGaussianDistribution:
    """sgm/modules/distributions/distributions.py: parameters = (mean | logvar) along the channel dim."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters   # fp32 [B, 2 C, h, w]
        self.deterministic = deterministic

    def sample(self, noise: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
        if self.deterministic:
            return self.mode(scale)
        if noise is None:
            b, c2, h, w = self.parameters.shape
            noise = torch.randn(b, c2 // 2, h, w, device=self.parameters.device)
        return ops.diag_gaussian(self.parameters, noise.float().contiguous(), scale)

    def mode(self, scale: float = 1.0) -> torch.Tensor:
        return ops.diag_gaussian(self.parameters, None, scale)


class AutoencoderKL(nn.Module, Packed):
    """sgm/models/autoencoder.py:282-316 (inference side: encoder, decoder, quant / post-quant 1x1 convolutions)."""

    def __init__(self, embed_dim: int, ddconfig: dict, ckpt_path=None, ignore_keys=(), lossconfig=None, monitor=None,
                 **kwargs):
        super().__init__()
        assert ddconfig["double_z"]
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        if ckpt_path is not None:
            raise NotImplementedError("load checkpoints with load_state_dict; the keys are the reference's")

    def add_denoise_encoder(self) -> "AutoencoderKL":
        """SR_backbone.__init__ (models/SR_model.py:22): denoise_encoder = deepcopy(encoder)."""
        import copy

        self.denoise_encoder = copy.deepcopy(self.encoder)
        self.denoise_encoder.__dict__.pop("_pk_cache", None)
        return self

    def moments(self, x: torch.Tensor, encoder: Optional[nn.Module] = None) -> torch.Tensor:
        """quant_conv(encoder(x)) -> fp32 NCHW [B, 2 embed_dim, h, w]."""
        ops.require_cuda(x, "b200sr.vae.AutoencoderKL")
        enc = self.encoder if encoder is None else encoder
        h = enc.forward_nhwc(to_nhwc(x))
        w = self._pk("quant.w", (self.quant_conv.weight,), lambda t: t.detach().float().reshape(t.shape[0], -1).contiguous())
        b = self._pk("quant.b", (self.quant_conv.bias,), _F32)
        return ops.pointwise_small(h, w, b, out_nchw_f32=True)

    def encode(self, x):
        assert not self.training, f"{self.__class__.__name__} only supports inference currently"
        return DiagonalGaussianDistribution(self.moments(x))

    def decode(self, z, scale: float = 1.0, **decoder_kwargs):
        """z: fp32 NCHW latent; `scale` multiplies it first (1 / scale_factor, SR_model.py:82)."""
        ops.require_cuda(z, "b200sr.vae.AutoencoderKL")
        w = self._pk("pq.w", (self.post_quant_conv.weight,), lambda t: t.detach().float().reshape(t.shape[0], -1).contiguous())
        b = self._pk("pq.b", (self.post_quant_conv.bias,), _F32)
        zin = ops.nchw_to_nhwc_bf16(z.float().contiguous(), scale)
        return self.decoder.forward_nhwc(ops.pointwise_small(zin, w, b))


class AutoencoderKLInferenceWrapper(AutoencoderKL):
    """autoencoder.py:319-321."""

    def encode(self, x):
        return super().encode(x).sample()


class FirstStage:
    """The three first-stage calls of SR_backbone (models/SR_model.py:57-85) for the driver:
    encode(img) = encode_first_stage_with_denoise(img, use_sample=False); decode(z) = decode_first_stage(z)."""

    name = "b200sr.vae.AutoencoderKL (SDXL VAE architecture, random-init weights)"

    def __init__(self, vae: AutoencoderKL, scale_factor: float = 0.13025):
        self.vae, self.scale_factor = vae, scale_factor

    @torch.no_grad()
    def encode(self, img: torch.Tensor) -> torch.Tensor:
        enc = getattr(self.vae, "denoise_encoder", None)
        return DiagonalGaussianDistribution(self.vae.moments(img, enc)).mode(self.scale_factor)

    @torch.no_grad()
    def encode_sample(self, img: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """encode_first_stage: the plain encoder, a posterior sample."""
        return DiagonalGaussianDistribution(self.vae.moments(img)).sample(noise, self.scale_factor)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.vae.decode(z, 1.0 / self.scale_factor)
