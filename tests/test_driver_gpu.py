"""The infer_dir-style driver on the GPU (row f1): whole per-image pipeline (bicubic, SR3 stage 1, VAE encode, cached
stage-2 loop, VAE decode, wavelet colour fix, uint8 pack), determinism and sharding invariance."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_end_to_end_and_sharding():
    from b200sr import colorfix, modules, sr3, vae
    from b200sr.driver import RestorationPipeline, run_sharded
    from oracle import configs, weights

    wrapper = modules.build_stage2(configs.STAGE2_UNET_TEST, configs.STAGE2_CONTROL_TEST).eval()
    weights.fill_(wrapper.state_dict(), 0)
    wrapper = wrapper.cuda()
    ae = vae.AutoencoderKLInferenceWrapper(configs.VAE_EMBED_DIM, dict(configs.VAE_DDCONFIG)).add_denoise_encoder().eval()
    weights.fill_(ae.state_dict(), 0)
    ae = ae.cuda()
    net = sr3.UNet(**configs.SR3_UNET).eval()
    weights.fill_(net.state_dict(), 0)
    diff = sr3.GaussianDiffusion(net.cuda(), image_size=224, channels=3, conditional=True)
    diff.set_new_noise_schedule(dict(configs.SR3_SCHEDULE, schedule="linear", n_timestep=8), device="cuda")

    def make():
        return RestorationPipeline(wrapper, diff, first_stage=vae.FirstStage(ae), device="cuda", num_steps=12,
                                   color_fix=colorfix.wavelet_reconstruction)

    g = torch.Generator().manual_seed(3)
    images = [torch.rand(1, 3, 32, 32, generator=g) * 2 - 1 for _ in range(3)]      # x8 -> 256^2, 32^2 latent
    caps = [tuple({"crossattn": torch.randn(1, 77, 2048, generator=g), "vector": torch.randn(1, 2816, generator=g)}
                  for _ in range(2)) for _ in range(3)]
    pipe = make()
    whole = run_sharded(pipe, images, caps, 0, 1, seed=11, keep=True)
    r = whole["results"][0]
    assert r["stage1"].shape == (1, 3, 256, 256) and r["latent"].shape == (1, 4, 32, 32)
    assert r["image"].shape == (1, 3, 256, 256) and torch.isfinite(r["image"]).all()
    assert r["uint8"].shape == (256, 256, 3) and r["uint8"].dtype == torch.uint8
    assert len(r["trace"]) == 12 and r["trace"][0] == "miss" and all(m >= 1 for m in whole["misses"])
    assert set(pipe.timings) == {"stage1_s", "encode_s", "stage2_s", "decode_s"}
    # same seeds -> same bits, independent of how the list is sharded (image i is seeded with seed + i)
    again = run_sharded(make(), images, caps, 0, 1, seed=11, keep=True)
    assert all(torch.equal(a["uint8"], b["uint8"]) for a, b in zip(whole["results"], again["results"]))
    part = run_sharded(make(), images, caps, 1, 2, seed=11, keep=True)
    assert part["indices"] == [1] and torch.equal(part["results"][0]["uint8"], whole["results"][1]["uint8"])
    pipe.engine.close()
