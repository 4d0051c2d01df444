"""infer_dir-style driver (row f1) on CPU through the test double: data flow of one image (bicubic, first-stage
encode, cached stage-2 loop, decode, colour fix, uint8 pack), determinism, and that a result does not depend on how
the image list is sharded."""
import torch

from oracle import configs, weights


def test_driver_flow_and_sharding(monkeypatch):
    import ops_double
    from b200sr import colorfix, modules, ops, vae
    from b200sr.driver import RestorationPipeline, run_sharded
    from test_stage2_cpu import _shim_for

    ops_double.install(monkeypatch, ops)
    wrapper = modules.build_stage2(configs.STAGE2_UNET_TEST, configs.STAGE2_CONTROL_TEST).eval()
    weights.fill_(wrapper.state_dict(), 0)
    ae = vae.AutoencoderKLInferenceWrapper(configs.VAE_EMBED_DIM, dict(configs.VAE_DDCONFIG)).add_denoise_encoder().eval()
    weights.fill_(ae.state_dict(), 0)

    def make():
        return RestorationPipeline(_shim_for(wrapper), None, first_stage=vae.FirstStage(ae), device="cpu", num_steps=4,
                                   color_fix=colorfix.wavelet_reconstruction)

    g = torch.Generator().manual_seed(1)
    images = [torch.rand(1, 3, 8, 8, generator=g) * 2 - 1 for _ in range(3)]
    caps = [tuple({"crossattn": torch.randn(1, 77, 2048, generator=g), "vector": torch.randn(1, 2816, generator=g)}
                  for _ in range(2)) for _ in range(3)]
    whole = run_sharded(make(), images, caps, 0, 1, seed=5, keep=True)
    assert whole["indices"] == [0, 1, 2] and all(m >= 1 for m in whole["misses"])
    r0 = whole["results"][0]
    assert r0["stage1"].shape == (1, 3, 64, 64) and r0["latent"].shape == (1, 4, 8, 8)
    assert r0["image"].shape == (1, 3, 64, 64) and r0["image"].dtype == torch.float32
    assert r0["uint8"].shape == (64, 64, 3) and r0["uint8"].dtype == torch.uint8
    assert len(r0["trace"]) == 4 and r0["trace"][0] == "miss"
    # rank 1 of 2 gets image 1 only; same seed per image index -> same bits as in the unsharded run
    part = run_sharded(make(), images, caps, 1, 2, seed=5, keep=True)
    assert part["indices"] == [1]
    assert torch.equal(part["results"][0]["uint8"], whole["results"][1]["uint8"])
    assert torch.equal(part["results"][0]["latent"], whole["results"][1]["latent"])
