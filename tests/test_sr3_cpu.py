"""CPU suite for the SR3 stage-1 path: oracle vs the real reference's golden vectors, drop-in
state_dict keys, and module wiring through the ops test double."""
import json
import os

import pytest
import torch

from oracle import configs, inputs, sr3 as osr3, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "sr3_32.pt"), weights_only=False)


@pytest.fixture(scope="module")
def net():
    from b200sr import sr3

    n = sr3.UNet(**configs.SR3_UNET).eval()
    weights.fill_(n.state_dict(), 0)
    return n


def test_state_dict_keys_match_reference(net):
    with open(os.path.join(GOLDEN, "sr3_keys.json")) as f:
        ref = json.load(f)
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert mine == ref


def test_oracle_matches_reference_golden(golden, net):
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    t = sd["downs.1.res_block.block1.block.3.weight"].double()
    assert [t.sum().item(), t.abs().sum().item()] == pytest.approx(golden["weight_checksum"], rel=1e-9)
    cond, noises = inputs.sr3_inputs(size=golden["size"], seed=0, steps=50)
    with torch.no_grad():
        eps = osr3.unet(sd, "", torch.cat([cond, noises[0]], dim=1), golden["level"])
    assert (eps - golden["eps"]).abs().max().item() < 1e-4
    # the 50-step ancestral loop, with the reference's RNG stream
    torch.manual_seed(golden["loop_seed"])
    seq = [torch.randn(cond.shape) for _ in range(50)] + [torch.zeros_like(cond)]
    sched = osr3.Schedule(**configs.SR3_SCHEDULE)
    with torch.no_grad():
        sr = osr3.p_sample_loop(lambda x, l: osr3.unet(sd, "", x, l), sched, cond, seq)
    assert (sr - golden["sr"]).abs().max().item() < 1e-3


def test_module_wiring_against_golden(monkeypatch, golden, net):
    import ops_double
    from b200sr import ops, sr3

    ops_double.install(monkeypatch, ops)
    cond, noises = inputs.sr3_inputs(size=golden["size"], seed=0, steps=50)
    with torch.no_grad():
        eps = net._forward_impl(torch.cat([cond, noises[0]], dim=1), golden["level"])
    assert eps.dtype == torch.float32 and eps.shape == golden["eps"].shape
    assert rel_l2(eps, golden["eps"]) < 1e-2
    # sampling loop through the drop-in GaussianDiffusion (test double => CPU): PSNR >= 40 dB vs the reference
    diff = sr3.GaussianDiffusion(net._forward_impl, image_size=224, channels=3, conditional=True)
    diff.set_new_noise_schedule(dict(configs.SR3_SCHEDULE, schedule="linear"), device="cpu")
    torch.manual_seed(golden["loop_seed"])
    seq = [torch.randn(cond.shape) for _ in range(50)] + [torch.zeros_like(cond)]
    sr = diff.p_sample_loop(cond, continous=False, noises=seq)
    mse = ((sr - golden["sr"]) ** 2).mean().item()
    assert 10 * torch.log10(torch.tensor(4.0 / mse)).item() >= 40.0  # images live in [-1, 1]: peak-to-peak 2


def test_cpu_tensors_are_refused(net):
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 6, 32, 32), torch.zeros(1, 1))
