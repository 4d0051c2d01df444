"""GPU parity of the output-side kernels (wavelet colour fix, bicubic + uint8 pack) vs the reference's golden outputs
and vs the oracle at image sizes of the path."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_wavelet_and_u8_vs_reference_golden():
    from b200sr import colorfix

    g = torch.load(os.path.join(GOLDEN, "colorfix_48.pt"), weights_only=False)
    out = colorfix.wavelet_reconstruction(g["content"].cuda(), g["style"].cuda())
    assert torch.allclose(out.cpu(), g["reconstruction"], rtol=0, atol=2e-6)       # fp32 stencil: summation order only
    hi, lo = colorfix.wavelet_decomposition(g["content"].cuda())
    assert torch.allclose(hi.cpu(), g["high"], rtol=0, atol=2e-6) and torch.allclose(lo.cpu(), g["low"], rtol=0, atol=2e-6)
    rec = g["reconstruction"][0].cuda()
    for key, (h, w) in (("u8_same", (48, 64)), ("u8_up", (70, 100)), ("u8_down", (31, 40))):
        u8 = colorfix.tensor_to_uint8(rec, h, w).cpu()
        diff = (u8.int() - g[key].int()).abs()
        # integer output: identical except where the fp32 value sits within rounding of an integer boundary
        assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 2e-3, (key, diff.max().item())
    assert torch.equal(colorfix.tensor_to_uint8(rec, 48, 64).cpu(), g["u8_same"])   # identity resize: bit-exact


def test_wavelet_1024_vs_oracle():
    from b200sr import colorfix
    from oracle import colorfix as ocf

    gen = torch.Generator(device="cuda").manual_seed(3)
    content = torch.rand(1, 3, 1024, 1024, generator=gen, device="cuda") * 2 - 1
    style = torch.rand(1, 3, 1024, 1024, generator=gen, device="cuda") * 2 - 1
    out = colorfix.wavelet_reconstruction(content, style)
    ref = ocf.wavelet_reconstruction(content, style)
    assert torch.allclose(out, ref, rtol=0, atol=3e-6)
