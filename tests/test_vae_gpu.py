"""GPU parity of the first stage (SDXL VAE encode / decode) and of its new kernels (pytest -m gpu)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
bf16 = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def psnr(a, b):
    a, b = a.float(), b.float()
    peak = (b.max() - b.min()).item()
    return 10 * math.log10(peak * peak / ((a - b) ** 2).mean().item())


@pytest.fixture(scope="module")
def model():
    from b200sr import vae
    from oracle import configs, weights

    m = vae.AutoencoderKLInferenceWrapper(configs.VAE_EMBED_DIM, dict(configs.VAE_DDCONFIG)).add_denoise_encoder().eval()
    weights.fill_(m.state_dict(), 0)
    return m.cuda()


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 64, 64, 128, 128), (2, 32, 48, 256, 256), (1, 16, 16, 512, 512), (1, 128, 128, 64, 64)])
def test_conv3x3_stride2_pad_bottom_right(n, h, w, cin, cout):
    """model.py:70-88: F.pad(x, (0, 1, 0, 1)) + conv(stride 2, padding 0)."""
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, h, w, cin, generator=g, device="cuda").to(bf16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g, device="cuda") * (1.0 / (9 * cin)) ** 0.5)
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    y = ops.conv3x3(x, ops.pack_conv3x3(wt), b, stride=2, pad_lo=0)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), wt.to(bf16).float(), b, stride=2).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 5e-3


def test_pointwise_small_and_diag_gaussian():
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    for cin, cout in ((8, 8), (4, 4), (8, 4)):
        x = torch.randn(2, 16, 24, cin, generator=g, device="cuda").to(bf16)
        w = torch.randn(cout, cin, generator=g, device="cuda")
        b = torch.randn(cout, generator=g, device="cuda")
        ref = F.linear(x.float(), w, b) * 0.5
        y = ops.pointwise_small(x, w, b, scale=0.5)
        assert rel_l2(y, ref) < 5e-3
        yf = ops.pointwise_small(x, w, b, out_nchw_f32=True, scale=0.5)
        assert torch.allclose(yf, ref.permute(0, 3, 1, 2), rtol=1e-5, atol=1e-5)
    m = torch.randn(2, 8, 16, 16, generator=g, device="cuda") * 3
    m[:, 4:] *= 20          # exercises the logvar clamp
    noise = torch.randn(2, 4, 16, 16, generator=g, device="cuda")
    mean, logvar = m.chunk(2, 1)
    assert torch.allclose(ops.diag_gaussian(m, None, 0.13025), mean * 0.13025)
    ref = (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise) * 0.13025
    assert torch.allclose(ops.diag_gaussian(m, noise, 0.13025), ref, rtol=1e-5, atol=1e-6)


def test_vae_matches_reference_golden(model):
    from b200sr import vae

    golden = torch.load(os.path.join(GOLDEN, "vae_64.pt"), weights_only=False)
    fs = vae.FirstStage(model)
    z = fs.encode(golden["img"].cuda())
    x = fs.decode(golden["z"].cuda())
    print(f"vae 64^2: z rel-L2 {rel_l2(z.cpu(), golden['z']):.3e}; decoded rel-L2 {rel_l2(x.cpu(), golden['decoded']):.3e} "
          f"PSNR {psnr(x.cpu(), golden['decoded']):.1f} dB")
    # BASELINE.json gates images by PSNR >= 40 dB (the 1e-2 rel-L2 bound is stated for the per-step eps prediction);
    # rel-L2 here is bounded by the bf16 policy's own error, checked against torch autocast in the next test
    assert rel_l2(z.cpu(), golden["z"]) < 2.5e-2
    assert rel_l2(x.cpu(), golden["decoded"]) < 2.5e-2 and psnr(x.cpu(), golden["decoded"]) >= 40.0


@pytest.mark.parametrize("size", [512, 1024])
def test_vae_vs_fp32_oracle(model, size):
    """Encode + decode at the sizes of the path (1024^2 = 128^2 latent) vs the fp32 oracle on the same GPU.
    At 1024^2 the mid attention sees 16384 tokens: the chunked single-head attention path."""
    from b200sr import vae
    from oracle import vae as ovae

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    g = torch.Generator(device="cuda").manual_seed(size)
    img = torch.rand(1, 3, size, size, generator=g, device="cuda") * 2 - 1
    fs = vae.FirstStage(model)
    with torch.no_grad():
        z_ref = ovae.encode_with_denoise(sd, img)
        x_ref = ovae.decode_first_stage(sd, z_ref)
        with torch.autocast("cuda", torch.bfloat16):   # the reference's own ae_dtype = bf16 policy on stock torch ops
            z_ac = ovae.encode_with_denoise(sd, img).float()
            x_ac = ovae.decode_first_stage(sd, z_ref).float()
    z = fs.encode(img)
    x = fs.decode(z_ref)
    ez, ex, ez_ac, ex_ac = rel_l2(z, z_ref), rel_l2(x, x_ref), rel_l2(z_ac, z_ref), rel_l2(x_ac, x_ref)
    print(f"vae {size}^2: z rel-L2 {ez:.3e} (torch bf16 autocast {ez_ac:.3e}); decoded rel-L2 {ex:.3e} "
          f"(autocast {ex_ac:.3e}) PSNR {psnr(x, x_ref):.1f} dB (autocast {psnr(x_ac, x_ref):.1f} dB)")
    # images are gated by PSNR >= 40 dB (BASELINE.json); the relative error may not exceed what the reference's bf16
    # autocast policy itself costs on stock torch kernels by more than a quarter
    assert psnr(x, x_ref) >= 40.0
    assert ez <= max(1e-2, 1.25 * ez_ac) and ex <= max(1e-2, 1.25 * ex_ac)
