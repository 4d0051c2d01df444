"""BASELINE config 3 / 1 companions to bench.py (which measures config 2): whole-image numbers.

* stage 2: the full 50-step restoration loop on one 128^2 latent (1024^2 image), CFG batch 2, with the
  reference's first-block cache ("dynamic step acceleration", img_threshold 0.3) and without it;
* stage 1: SR3 x8 16^2 -> 128^2, 50 ancestral steps.
Synthetic inputs, seeded random weights (the hit / miss pattern of the cache therefore belongs to these
weights, not to the released checkpoint).  Prints one JSON object; also written to gpurun_out/image_bench.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import modules, ops, sr3
from b200sr.sampling import Stage2Engine
from oracle import configs, inputs, weights

dev = torch.device("cuda", 0)
res = {}

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3

# ---- stage 2 ----
wrapper = modules.build_stage2(configs.STAGE2_UNET, configs.STAGE2_CONTROL).eval()
weights.fill_(wrapper.state_dict(), 0)
wrapper = wrapper.to(dev)
x0, c, uc = inputs.stage2_inputs(latent=128, seed=1234)
eng = Stage2Engine(wrapper, device=dev)
eng.set_condition({k: v.to(dev) for k, v in c.items()}, {k: v.to(dev) for k, v in uc.items()})
g = torch.Generator(device="cpu").manual_seed(7)
z0 = torch.randn(x0.shape, generator=g).to(dev)
noises = [torch.randn(x0.shape, generator=g).to(dev) for _ in range(eng.sched.num_steps)]
for name, thr in (("uncached", 0.0), ("first_block_cache_thr0.3", 0.3)):
    t = timed(lambda: eng.sample(z0, noises, threshold=thr), 2)
    trace = "".join("H" if kind == "hit" else "M" for kind, _ in eng.trace)
    res["stage2_" + name] = {"s_per_image": t, "images_per_s": 1.0 / t, "steps": eng.sched.num_steps, "trace": trace}
    print(name, f"{t:.3f} s/image", trace, flush=True)
del eng, wrapper
torch.cuda.empty_cache()

# ---- stage 1 (SR3) ----
net = sr3.UNet(**configs.SR3_UNET).eval()
weights.fill_(net.state_dict(), 0)
net = net.to(dev)
diff = sr3.GaussianDiffusion(net, image_size=224, channels=3, conditional=True)
diff.set_new_noise_schedule(dict(configs.SR3_SCHEDULE, schedule="linear"), device="cuda")
cond, _ = inputs.sr3_inputs(size=128, seed=0, steps=1)
cond = cond.to(dev)
seq = [torch.randn(cond.shape, device=dev) for _ in range(50)] + [torch.zeros_like(cond)]
for name, graphs in (("eager", False), ("graph", True)):
    diff.use_graphs = graphs
    t = timed(lambda: diff.p_sample_loop(cond, continous=False, noises=seq), 3)
    res["sr3_x8_16_to_128_" + name] = {"s_per_image": t, "images_per_s": 1.0 / t, "steps": configs.SR3_SCHEDULE["n_timestep"]}
    print("sr3", name, f"{t:.4f} s/image", flush=True)
# ---- stage 1 at the x4 size of config 3: 256^2 -> 1024^2, 50 steps ----
g4 = torch.Generator().manual_seed(5)
lr = torch.rand(1, 3, 256, 256, generator=g4) * 2 - 1
cond4 = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False).clamp(-1, 1).to(dev)
diff4 = sr3.GaussianDiffusion(net, image_size=1024, channels=3, conditional=True)
diff4.set_new_noise_schedule(dict(configs.SR3_SCHEDULE, schedule="linear"), device="cuda")
t = timed(lambda: diff4.p_sample_loop(cond4, continous=False), 1)
res["sr3_x4_256_to_1024_graph"] = {"s_per_image": t, "images_per_s": 1.0 / t, "steps": configs.SR3_SCHEDULE["n_timestep"]}
print("sr3 x4 1024", f"{t:.3f} s/image", flush=True)
two = t + res["stage2_first_block_cache_thr0.3"]["s_per_image"]
res["two_stage_x4_1024_cached"] = {"s_per_image": two, "images_per_s": 1.0 / two,
                                   "note": "SR3 50 steps at 1024^2 + 50 cached stage-2 steps; VAE / captioner out of scope (synthetic latents)"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "image_bench.json"), "w"), indent=1)
print(json.dumps(res))
