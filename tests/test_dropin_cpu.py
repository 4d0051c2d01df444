"""The drop-in claim, shown: the reference's OWN sampler stack — DiscreteDenoiserWithControl.__call__
(sgm/modules/diffusionmodules/denoiser.py:66-78), RestoreEDMSampler.step / denoise (sampling.py:548-694), LinearCFG
(guiders.py:44-74) and the first-block cache context (models/modules/DFBCache.py) — imported unmodified from
/root/reference, drives ``b200sr.modules.ControlWrapper`` exactly as it drives its own ``ControlWrapper``:
positional ``network(input * c_in, c_noise, cond, control_scale, fbcache_mode, partial_info)``, the ``"h"`` entry of
the returned partial_info read by the sampler, fp32 eps back.  The kernels are replaced by the CPU test double
(tests/ops_double.py); the module wiring, signatures and return conventions are the product's.

Skipped where /root/reference does not exist (the GPU box)."""
import math
import os

import pytest
import torch

from oracle import configs, inputs, reference_import, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.skipif(not reference_import.available(), reason="needs the reference checkout")


def test_reference_sampler_drives_b200sr_wrapper(monkeypatch):
    import ops_double
    from b200sr import modules, ops

    ops_double.install(monkeypatch, ops)
    golden = torch.load(os.path.join(GOLDEN, "stage2_test_16.pt"), weights_only=False)
    wrapper = modules.build_stage2(configs.STAGE2_UNET_TEST, configs.STAGE2_CONTROL_TEST).eval()
    weights.fill_(wrapper.state_dict(), 0)
    den, smp = reference_import.stage2_sampler(device="cpu")
    from models.modules.DFBCache import MyCacheContext, cache_context

    calls = []

    def denoiser(inp, sigma, cc, control_scale=1.0, fbcache_mode="none", partial_info=None):
        calls.append(fbcache_mode)
        return den(wrapper, inp, sigma, cc, control_scale, fbcache_mode, partial_info)   # denoiser.py:66-78

    _, c, uc = inputs.stage2_inputs(latent=golden["latent"], seed=1234)
    z, s_in, sigmas, _, c, uc = smp.init_loop(golden["z0"].clone(), c, uc, num_steps=50)
    assert torch.equal(sigmas, golden["sigmas"])
    thr, trace = golden["threshold"], []
    with torch.no_grad(), cache_context(MyCacheContext()):
        for i in range(golden["steps"]):
            torch.manual_seed(1000 + i)                      # the reference draws randn_like(x) from the global RNG
            z, new_thr = smp.step(z, i, s_in, sigmas, denoiser, c, uc, x_center=None, control_scale=1.0, threshold=thr)
            trace.append("hit" if (new_thr == thr and i > 0) else "miss")
            thr = new_thr
    assert trace == [t[0] for t in golden["trace"]]
    # every step asked for stage 1; only the misses went on to stage 2 (SR_modules.py:660-730 protocol)
    assert calls.count("input_stage1") == golden["steps"] and calls.count("input_stage2") == trace.count("miss")
    assert z.dtype == torch.float32
    mse = ((z - golden["z_final"]) ** 2).mean().item()
    peak = (golden["z_final"].max() - golden["z_final"].min()).item()
    assert 10 * math.log10(peak * peak / mse) >= 40.0


def test_reference_yaml_targets_resolve_to_b200sr_classes():
    """the swap is a change of `target:` strings; the constructors take the YAML's params."""
    import yaml
    from b200sr import modules

    with open(os.path.join(reference_import.REFERENCE_ROOT, "model_configs", "juggernautXL.yaml")) as f:
        params = yaml.safe_load(f)["model"]["params"]
    for key, cls in (("control_stage_config", modules.GLVControl), ("network_config", modules.LightGLVUNet)):
        with torch.device("meta"):
            m = cls(**params[key]["params"])
        assert isinstance(m, modules.UNetModel)
