"""First stage (SDXL VAE) on CPU: the oracle against the reference's golden outputs,
the reference's state_dict keys, and the drop-in modules' host wiring through the test double."""
import json
import math
import os

import pytest
import torch

from oracle import configs, vae as ovae, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def psnr(a, b):
    a, b = a.float(), b.float()
    peak = (b.max() - b.min()).item()
    return 10 * math.log10(peak * peak / ((a - b) ** 2).mean().item())


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "vae_64.pt"), weights_only=False)


@pytest.fixture(scope="module")
def model():
    from b200sr import vae

    m = vae.AutoencoderKLInferenceWrapper(configs.VAE_EMBED_DIM, dict(configs.VAE_DDCONFIG)).add_denoise_encoder().eval()
    weights.fill_(m.state_dict(), 0)
    return m


def test_state_dict_keys_match_reference(model):
    with open(os.path.join(GOLDEN, "vae_keys.json")) as f:
        ref = json.load(f)
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert ours == ref


def test_oracle_matches_reference_golden(golden, model):
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    assert [sd["decoder.mid.block_1.conv1.weight"].double().sum().item(),
            sd["decoder.mid.block_1.conv1.weight"].double().abs().sum().item()] == pytest.approx(golden["weight_checksum"])
    with torch.no_grad():
        m = ovae.moments(sd, golden["img"], "denoise_encoder.")
        z = ovae.encode_with_denoise(sd, golden["img"])
        x = ovae.decode_first_stage(sd, golden["z"])
    assert (m - golden["moments"]).abs().max().item() < 1e-5
    assert (z - golden["z"]).abs().max().item() < 1e-5
    assert (x - golden["decoded"]).abs().max().item() < 1e-4


def test_module_wiring_against_golden(monkeypatch, golden, model):
    """Encoder / decoder modules driven through the CPU double == the reference's fp32 outputs within bf16 tolerance."""
    import ops_double
    from b200sr import ops, vae

    ops_double.install(monkeypatch, ops)
    fs = vae.FirstStage(model)
    with torch.no_grad():
        m = model.moments(golden["img"], model.denoise_encoder)
        z = fs.encode(golden["img"])
        x = fs.decode(golden["z"])
        sample = fs.encode_sample(golden["img"], torch.zeros(1, 4, 8, 8))
    assert m.dtype == torch.float32 and m.shape == golden["moments"].shape
    assert rel_l2(m, golden["moments"]) < 2e-2
    assert rel_l2(z, golden["z"]) < 2e-2
    assert x.dtype == torch.float32 and x.shape == golden["decoded"].shape
    assert rel_l2(x, golden["decoded"]) < 2e-2 and psnr(x, golden["decoded"]) >= 35.0
    assert sample.shape == z.shape   # zero noise: the plain encoder's mean * scale_factor
