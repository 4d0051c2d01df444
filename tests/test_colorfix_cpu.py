"""Output side (row f3) on CPU: oracle vs the reference's golden outputs; host module through the test double."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_matches_reference_golden():
    from oracle import colorfix as ocf

    g = torch.load(os.path.join(GOLDEN, "colorfix_48.pt"), weights_only=False)
    assert torch.equal(ocf.wavelet_reconstruction(g["content"], g["style"]), g["reconstruction"])
    hi, lo = ocf.wavelet_decomposition(g["content"])
    assert torch.equal(hi, g["high"]) and torch.equal(lo, g["low"])
    assert torch.equal(ocf.tensor_to_uint8(g["reconstruction"][0], 48, 64), g["u8_same"])


def test_host_module_through_double(monkeypatch):
    import ops_double
    from b200sr import colorfix, ops

    ops_double.install(monkeypatch, ops)
    g = torch.load(os.path.join(GOLDEN, "colorfix_48.pt"), weights_only=False)
    out = colorfix.wavelet_reconstruction(g["content"], g["style"])
    assert torch.allclose(out, g["reconstruction"], rtol=0, atol=1e-6)
    assert torch.equal(colorfix.tensor_to_uint8(g["reconstruction"][0], 48, 64), g["u8_same"])
