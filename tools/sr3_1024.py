"""BASELINE config 3, stage 1 at full size: one SR3 UNet call on a 1024^2 image (x4 of a 256^2 input) against the
fp32 oracle evaluated on the GPU, and the time of one graphed ancestral step."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import sr3
from oracle import configs, weights, sr3 as osr3
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = sr3.UNet(**configs.SR3_UNET).eval(); weights.fill_(net.state_dict(), 0); net = net.cuda()
g = torch.Generator().manual_seed(5)
lr = torch.rand(1, 3, S // 4, S // 4, generator=g) * 2 - 1
cond = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False).clamp(-1, 1).cuda()
x = torch.randn(1, 3, S, S, generator=g).cuda()
level = torch.tensor([[0.42]], device="cuda")
inp = torch.cat([cond, x], 1)
sd = {k: v.detach() for k, v in net.state_dict().items()}
with torch.no_grad():
    eps = net(inp, level)
    ref = osr3.unet(sd, "", inp, level)
torch.cuda.synchronize()
res = {"size": S}
if ref is not None:
    res["eps_rel_l2_vs_fp32_oracle"] = float(((eps - ref).norm() / ref.norm()))
diff = sr3.GaussianDiffusion(net, image_size=S, channels=3, conditional=True)
diff.set_new_noise_schedule(dict(configs.SR3_SCHEDULE, schedule="linear"), device="cuda")
nz = torch.randn_like(x)
for _ in range(3): diff._p_sample_graphed(x, 10, condition_x=cond, noise=nz)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): diff._p_sample_graphed(x, 10, condition_x=cond, noise=nz)
e1.record(); torch.cuda.synchronize()
res["ms_per_step_graphed"] = e0.elapsed_time(e1) / 5
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sr3_{S}.json"), "w"))
# per-launch breakdown of one eager step (CUDA events around every instrumented op, GPU kept busy while the host enqueues)
from b200sr import ops
recs = []
torch.cuda.synchronize(); torch.cuda._sleep(int(0.3 * 1.9e9))
ops.set_profile(recs)
with torch.no_grad():
    net(inp, level)
ops.set_profile(None); torch.cuda.synchronize()
agg = {}
for kind, work, a, b, desc in recs:
    d = agg.setdefault((kind, desc), [0.0, 0.0, 0]); d[0] += work; d[1] += a.elapsed_time(b); d[2] += 1
tot = sum(v[1] for v in agg.values())
print(f"event-timed total {tot:.2f} ms over {sum(v[2] for v in agg.values())} launches")
for (kind, desc), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    rate = v[0] / (v[1] * 1e-3) / 1e12 if v[1] else 0
    print(f"  {kind:11s} {desc:38s} n={v[2]:3d} {v[1]:7.3f} ms  {rate:8.1f} {'TB/s' if kind in ('group_norm','layer_norm') else 'TF/s'}")
