"""Drop-in SR3 stage-1 denoiser: the conditional UNet and the DDPM ancestral sampling loop, with the
reference's constructor arguments, ``forward`` signatures and ``state_dict`` keys, computing on the
sm_100a kernels.

Mirrored reference code (relative to the reference root):
  PositionalEncoding / FeatureWiseAffine / Swish / Upsample / Downsample / Block / ResnetBlock /
  SelfAttention / ResnetBlocWithAttn / UNet        models/sr3_model/sr3_modules/unet.py:19-261
  GaussianDiffusion (schedule, p_sample loop)      models/sr3_model/sr3_modules/diffusion.py:65-250

Tensor conventions are those of ``b200sr.modules`` (NCHW-shaped, channels-last bf16 between modules;
``UNet.forward`` takes / returns fp32 NCHW like the reference).  The single-head attention
(width C = 512, 1/sqrt(C) scale, full softmax) is three projection GEMMs, one fp32 score GEMM,
a row-softmax kernel and one P @ V GEMM against a transposed-V projection.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .modules import Packed, _F32, _gn, _linear, from_nhwc, to_nhwc, tokens_bf16

bf16 = torch.bfloat16


def exists(x):
    return x is not None


class PositionalEncoding(nn.Module):
    """unet.py:19-32 — sin | cos of noise_level * exp(-ln(1e4) k / count)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, noise_level):
        # noise_level: [B, 1] -> bf16 [B, dim]
        return ops.sinusoid_embedding(noise_level.reshape(-1), self.dim, 1e4, sin_first=True)


class Swish(nn.Module):
    def forward(self, x):  # only reached if someone calls it stand-alone
        return ops.silu(tokens_bf16(x))


class FeatureWiseAffine(nn.Module, Packed):
    """unet.py:35-51 (use_affine_level=False: additive).  The add itself is fused into block1's conv epilogue."""

    def __init__(self, in_channels, out_channels, use_affine_level=False):
        super().__init__()
        assert not use_affine_level
        self.use_affine_level = use_affine_level
        self.noise_func = nn.Sequential(nn.Linear(in_channels, out_channels))

    def shift(self, noise_embed):
        """fp32 [B, out_channels] to add per image."""
        return _linear(self, "nf", self.noise_func[0], noise_embed, out_fp32=True)


def _conv_any(holder: Packed, name: str, conv: nn.Conv2d, x_nhwc, **kw):
    """3x3 conv for any Cin: Cin % 64 == 0 directly, otherwise zero-padded up to the next multiple of 64 (the 6-channel
    stem; the 96 / 160-channel concatenations of an inner_channel = 32 network)."""
    cin = conv.weight.shape[1]
    b = holder._pk(name + ".b", (conv.bias,), _F32) if conv.bias is not None else None
    if cin % 64 == 0:
        w = holder._pk(name + ".w", (conv.weight,), ops.pack_conv3x3)
        return ops.conv3x3(x_nhwc, w, b, **kw)
    cpad = (cin + 63) // 64 * 64
    w = holder._pk(name + ".wpad", (conv.weight,), lambda t: ops.pack_conv3x3_padded(t, cpad))
    return ops.conv3x3(ops.pad_channels(x_nhwc, cpad), w, b, **kw)


class Upsample(nn.Module, Packed):
    """unet.py:59-66."""

    def __init__(self, dim):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.conv = nn.Conv2d(dim, dim, 3, padding=1)

    def forward_nhwc(self, x):
        return _conv_any(self, "conv", self.conv, ops.upsample2x(x))

    def forward(self, x):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class Downsample(nn.Module, Packed):
    """unet.py:69-75."""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, 2, 1)

    def forward_nhwc(self, x):
        return _conv_any(self, "conv", self.conv, x, stride=2)

    def forward(self, x):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class Block(nn.Module, Packed):
    """unet.py:81-92 — GroupNorm -> Swish -> (Dropout: identity at inference) -> conv3x3."""

    def __init__(self, dim, dim_out, groups=32, dropout=0):
        super().__init__()
        self.block = nn.Sequential(nn.GroupNorm(groups, dim), Swish(),
                                   nn.Dropout(dropout) if dropout != 0 else nn.Identity(),
                                   nn.Conv2d(dim, dim_out, 3, padding=1))

    def forward_nhwc(self, x, **kw):
        from .modules import _gn_silu_conv3x3

        if x.shape[-1] % 64 == 0:   # GroupNorm + Swish folded into the convolution's input tiles where it applies
            return _gn_silu_conv3x3(self, "conv", self.block[0], self.block[3], x, **kw)
        return _conv_any(self, "conv", self.block[3], _gn(self.block[0], x, silu=True), **kw)

    def forward(self, x):
        return from_nhwc(self.forward_nhwc(to_nhwc(x)))


class ResnetBlock(nn.Module, Packed):
    """unet.py:95-111."""

    def __init__(self, dim, dim_out, noise_level_emb_dim=None, dropout=0, use_affine_level=False, norm_groups=32):
        super().__init__()
        self.noise_func = FeatureWiseAffine(noise_level_emb_dim, dim_out, use_affine_level)
        self.block1 = Block(dim, dim_out, groups=norm_groups)
        self.block2 = Block(dim_out, dim_out, groups=norm_groups, dropout=dropout)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward_nhwc(self, x, time_emb):
        h = self.block1.forward_nhwc(x, rowvec=self.noise_func.shift(time_emb))
        res = x if isinstance(self.res_conv, nn.Identity) else _linear(self, "res", self.res_conv, x)
        return self.block2.forward_nhwc(h, residual=res)

    def forward(self, x, time_emb):
        return from_nhwc(self.forward_nhwc(to_nhwc(x), tokens_bf16(time_emb)))


class SelfAttention(nn.Module, Packed):
    """unet.py:114-143 — one head of width C, scale 1/sqrt(C), softmax over all H*W keys."""

    def __init__(self, in_channel, n_head=1, norm_groups=32):
        super().__init__()
        assert n_head == 1, "the SR3 configuration uses a single head"
        self.n_head = n_head
        self.norm = nn.GroupNorm(norm_groups, in_channel)
        self.qkv = nn.Conv2d(in_channel, in_channel * 3, 1, bias=False)
        self.out = nn.Conv2d(in_channel, in_channel, 1)

    def forward_nhwc(self, x):
        b, h, w, c = x.shape
        n = _gn(self.norm, x).view(b, h * w, c)
        wq, wk, wv = self._pk("qkv", (self.qkv.weight,),
                              lambda t: tuple(p.contiguous() for p in ops.pack_linear(t).chunk(3, dim=0)))
        outs = []
        t = h * w
        tp = max(8, (t + 7) // 8 * 8)                              # GEMM dims must be multiples of 8
        for i in range(b):
            tok = n[i]                                             # [HW, C]
            if tp != t:                                            # tiny feature maps only: zero-pad the tokens
                tok = torch.zeros(tp, c, dtype=bf16, device=x.device)
                tok[:t].copy_(n[i])
            q, k = ops.gemm(tok, wq), ops.gemm(tok, wk)            # [T, C]
            v_t = ops.gemm(wv, tok, w_dynamic=True)                          # [C, T]  (V transposed: B operand of P @ V)
            o_i = ops.single_head_attention(q, k, v_t, 1.0 / math.sqrt(c), valid_keys=t)   # query chunks, never T x T
            outs.append(o_i if tp == t else o_i[:t].contiguous())
        o = outs[0].unsqueeze(0) if b == 1 else torch.stack(outs, 0)
        y = _linear(self, "out", self.out, o.reshape(b, h * w, c), residual=x.view(b, h * w, c))
        return y.view(b, h, w, c)

    def forward(self, input):
        return from_nhwc(self.forward_nhwc(to_nhwc(input)))


class ResnetBlocWithAttn(nn.Module):
    """unet.py:146-159."""

    def __init__(self, dim, dim_out, *, noise_level_emb_dim=None, norm_groups=32, dropout=0, with_attn=False):
        super().__init__()
        self.with_attn = with_attn
        self.res_block = ResnetBlock(dim, dim_out, noise_level_emb_dim, norm_groups=norm_groups, dropout=dropout)
        if with_attn:
            self.attn = SelfAttention(dim_out, norm_groups=norm_groups)

    def forward_nhwc(self, x, time_emb):
        x = self.res_block.forward_nhwc(x, time_emb)
        return self.attn.forward_nhwc(x) if self.with_attn else x

    def forward(self, x, time_emb):
        return from_nhwc(self.forward_nhwc(to_nhwc(x), tokens_bf16(time_emb)))


class UNet(nn.Module, Packed):
    """unet.py:162-261."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                 attn_res=(8), res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=128):
        super().__init__()
        assert with_noise_level_emb
        attn_res = attn_res if isinstance(attn_res, (list, tuple)) else [attn_res]
        noise_level_channel = inner_channel
        self.noise_level_mlp = nn.Sequential(PositionalEncoding(inner_channel),
                                             nn.Linear(inner_channel, inner_channel * 4), Swish(),
                                             nn.Linear(inner_channel * 4, inner_channel))
        num_mults = len(channel_mults)
        pre_channel = inner_channel
        feat_channels = [pre_channel]
        now_res = image_size
        downs = [nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)]
        for ind in range(num_mults):
            is_last = ind == num_mults - 1
            use_attn = now_res in attn_res
            channel_mult = inner_channel * channel_mults[ind]
            for _ in range(0, res_blocks):
                downs.append(ResnetBlocWithAttn(pre_channel, channel_mult, noise_level_emb_dim=noise_level_channel,
                                                norm_groups=norm_groups, dropout=dropout, with_attn=use_attn))
                feat_channels.append(channel_mult)
                pre_channel = channel_mult
            if not is_last:
                downs.append(Downsample(pre_channel))
                feat_channels.append(pre_channel)
                now_res = now_res // 2
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([
            ResnetBlocWithAttn(pre_channel, pre_channel, noise_level_emb_dim=noise_level_channel,
                               norm_groups=norm_groups, dropout=dropout, with_attn=True),
            ResnetBlocWithAttn(pre_channel, pre_channel, noise_level_emb_dim=noise_level_channel,
                               norm_groups=norm_groups, dropout=dropout, with_attn=False)])
        ups = []
        for ind in reversed(range(num_mults)):
            is_last = ind < 1
            use_attn = now_res in attn_res
            channel_mult = inner_channel * channel_mults[ind]
            for _ in range(0, res_blocks + 1):
                ups.append(ResnetBlocWithAttn(pre_channel + feat_channels.pop(), channel_mult,
                                              noise_level_emb_dim=noise_level_channel, norm_groups=norm_groups,
                                              dropout=dropout, with_attn=use_attn))
                pre_channel = channel_mult
            if not is_last:
                ups.append(Upsample(pre_channel))
                now_res = now_res * 2
        self.ups = nn.ModuleList(ups)
        self.final_conv = Block(pre_channel, out_channel if exists(out_channel) else in_channel, groups=norm_groups)

    def _time(self, time):
        enc = self.noise_level_mlp[0](time.to(torch.float32))
        h = _linear(self, "t1", self.noise_level_mlp[1], enc, act=1)
        return _linear(self, "t3", self.noise_level_mlp[3], h)

    def forward(self, x, time):
        ops.require_cuda(x, "b200sr.sr3.UNet")
        return self._forward_impl(x, time)

    def _forward_impl(self, x, time):
        t = self._time(time)
        h = to_nhwc(x)
        feats = []
        for layer in self.downs:
            if isinstance(layer, ResnetBlocWithAttn):
                h = layer.forward_nhwc(h, t)
            elif isinstance(layer, nn.Conv2d):
                h = _conv_any(self, "stem", layer, h)
            else:
                h = layer.forward_nhwc(h)
            feats.append(h)
        for layer in self.mid:
            h = layer.forward_nhwc(h, t)
        for layer in self.ups:
            if isinstance(layer, ResnetBlocWithAttn):
                h = layer.forward_nhwc(ops.concat_add(h, feats.pop()), t)
            else:
                h = layer.forward_nhwc(h)
        # final_conv: GN + Swish + conv to <= 4 channels, written straight to fp32 NCHW
        blk = self.final_conv.block
        g = _gn(blk[0], h, silu=True)
        w8, b8 = self._pk("final.wb8", (blk[3].weight, blk[3].bias), ops.pack_conv3x3_few_out)
        return ops.conv3x3_to_nchw_f32(g, w8, b8, blk[3].weight.shape[0])


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    """diffusion.py:20-56 (the schedules the shipped configs use)."""
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    if schedule == "quad":
        return np.linspace(linear_start**0.5, linear_end**0.5, n_timestep, dtype=np.float64) ** 2
    if schedule == "const":
        return linear_end * np.ones(n_timestep, dtype=np.float64)
    raise NotImplementedError(schedule)


class GaussianDiffusion(nn.Module):
    """diffusion.py:65-250, inference side: schedule tables + ancestral sampling loop.  Per-step scalars live
    in one device table; the x0-prediction / clamp / posterior-mean / noise update is one kernel."""

    def __init__(self, denoise_fn, image_size, channels=3, loss_type="l1", conditional=True, schedule_opt=None):
        super().__init__()
        self.channels, self.image_size, self.denoise_fn = channels, image_size, denoise_fn
        self.loss_type, self.conditional = loss_type, conditional

    def set_new_noise_schedule(self, schedule_opt, device):
        betas = make_beta_schedule(schedule=schedule_opt["schedule"], n_timestep=schedule_opt["n_timestep"],
                                   linear_start=schedule_opt["linear_start"], linear_end=schedule_opt["linear_end"])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.sqrt_alphas_cumprod_prev = np.sqrt(np.append(1.0, ac))
        self.num_timesteps = int(betas.shape[0])
        f32 = lambda a: torch.tensor(a, dtype=torch.float32, device=device)  # noqa: E731
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(ac))
        self.register_buffer("alphas_cumprod_prev", f32(ac_prev))
        self.register_buffer("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1.0 - ac)))
        self.register_buffer("log_one_minus_alphas_cumprod", f32(np.log(1.0 - ac)))
        self.register_buffer("sqrt_recip_alphas_cumprod", f32(np.sqrt(1.0 / ac)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1.0 / ac - 1)))
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.register_buffer("posterior_variance", f32(post_var))
        self.register_buffer("posterior_log_variance_clipped", f32(np.log(np.maximum(post_var, 1e-20))))
        self.register_buffer("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1.0 - ac)))
        self.register_buffer("posterior_mean_coef2", f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)))
        # [T, 5] per-step scalars of the update kernel and [T] continuous noise levels
        self._step_table = torch.stack([self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod,
                                        self.posterior_mean_coef1, self.posterior_mean_coef2,
                                        self.posterior_log_variance_clipped], dim=1).contiguous()
        self._levels = f32(self.sqrt_alphas_cumprod_prev[1:].astype(np.float32))

    @torch.no_grad()
    def p_sample(self, x, t, clip_denoised=True, condition_x=None, noise=None):
        """diffusion.py:152-176."""
        assert clip_denoised
        level = self._levels[t].reshape(1, 1).repeat(x.shape[0], 1)
        inp = torch.cat([condition_x, x], dim=1) if condition_x is not None else x
        eps = self.denoise_fn(inp, level)
        if t > 0 and noise is None:
            noise = torch.randn_like(x)
        return ops.sr3_update(x, eps, noise if t > 0 else None, self._step_table[t])

    use_graphs = True  # one CUDA graph per (shape): ~150 launches of a step replayed in one go
    max_cached_shapes = 4  # least-recently-used input shapes beyond this are dropped (infer_dir runs see many sizes)

    @torch.no_grad()
    def _p_sample_graphed(self, x, t, condition_x=None, noise=None):
        """p_sample through a captured CUDA graph: the per-step scalars, the noise level and the noise are copied into
        static buffers, then the whole step (concat, UNet, update kernel) replays.  At t = 0 the noise buffer is zero,
        which equals the reference's `noise = zeros` branch (diffusion.py:174)."""
        key = (tuple(x.shape), None if condition_x is None else tuple(condition_x.shape))
        table = self.__dict__.setdefault("_graph_state", {})
        st = table.pop(key, None)
        if st is not None:
            table[key] = st                      # most recently used last
        if st is None:
            while len(table) >= self.max_cached_shapes:   # each entry pins a CUDA graph and its private memory pool
                table.pop(next(iter(table)))
            st = table[key] = {"x": torch.empty_like(x), "noise": torch.zeros_like(x),
                               "cond": None if condition_x is None else torch.empty_like(condition_x),
                               "level": torch.empty(x.shape[0], 1, dtype=torch.float32, device=x.device),
                               "scal": torch.empty(self._step_table.shape[1], dtype=torch.float32, device=x.device),
                               "graph": None, "cond_src": None}
        st["x"].copy_(x)
        src = None if condition_x is None else (condition_x, condition_x._version)
        if condition_x is not None and (st["cond_src"] is None or st["cond_src"][0] is not condition_x
                                        or st["cond_src"][1] != condition_x._version):
            st["cond"].copy_(condition_x)   # also when the same tensor object was refilled in place
            st["cond_src"] = src
        if t > 0:
            st["noise"].copy_(noise) if noise is not None else st["noise"].normal_()
        else:
            st["noise"].zero_()
        st["level"].copy_(self._levels[t].reshape(1, 1).expand(x.shape[0], 1))
        st["scal"].copy_(self._step_table[t])

        def body():
            inp = torch.cat([st["cond"], st["x"]], dim=1) if st["cond"] is not None else st["x"]
            st["out"] = ops.sr3_update(st["x"], self.denoise_fn(inp, st["level"]), st["noise"], st["scal"])

        if st["graph"] is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()  # eager warm-up: packs the weights, sizes the workspaces
                body()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            st["graph"] = g
        st["graph"].replay()
        return st["out"].clone()

    @torch.no_grad()
    def p_sample_loop(self, x_in, continous=False, noises: Optional[List[torch.Tensor]] = None):
        """diffusion.py:178-201.  `noises` (optional) = [initial image, noise of step 0, 1, ...] for reproducible runs."""
        device = self.betas.device
        sample_inter = 1 | (self.num_timesteps // 10)
        cond = x_in if self.conditional else None
        shape = x_in.shape if self.conditional else x_in
        img = noises[0].to(device) if noises is not None else torch.randn(shape, device=device)
        ret_img = x_in if self.conditional else img
        step = self._p_sample_graphed if (self.use_graphs and img.is_cuda) else self.p_sample
        for k, i in enumerate(reversed(range(0, self.num_timesteps))):
            img = step(img, i, condition_x=cond, noise=None if noises is None else noises[1 + k].to(device))
            if i % sample_inter == 0:
                ret_img = torch.cat([ret_img, img], dim=0)
        return ret_img if continous else ret_img[-1]

    @torch.no_grad()
    def sample(self, batch_size=1, continous=False):
        return self.p_sample_loop((batch_size, self.channels, self.image_size, self.image_size), continous)

    @torch.no_grad()
    def super_resolution(self, x_in, continous=False):
        return self.p_sample_loop(x_in, continous)

    def forward(self, x, *args, **kwargs):
        raise NotImplementedError("training (p_losses) is outside the B200 inference hot path")
