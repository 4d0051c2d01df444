"""Per-shape timing of the dense kernels over one eager stage-2 step (CUDA events per launch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import modules, ops
from b200sr.sampling import Stage2Engine
from oracle import configs, inputs, weights

dev = torch.device("cuda")
w = modules.build_stage2(configs.STAGE2_UNET, configs.STAGE2_CONTROL).eval()
if "--fill" in sys.argv:
    weights.fill_(w.state_dict(), 0)
w = w.to(dev)
x0, c, uc = inputs.stage2_inputs(latent=128, seed=1234)
eng = Stage2Engine(w, use_graphs=False, device=dev)
eng.set_condition({k: v.to(dev) for k, v in c.items()}, {k: v.to(dev) for k, v in uc.items()})
x = x0.to(dev); noise = torch.randn_like(x)
for _ in range(2):
    eng.step(x, 3, noise, 0.0)
recs = []
ops.set_profile(recs)
for _ in range(3):
    eng.step(x, 3, noise, 0.0)
ops.set_profile(None)
torch.cuda.synchronize()
agg = {}
for kind, fl, a, b, desc in recs:
    d = agg.setdefault((kind, desc), [0.0, 0.0, 0])
    d[0] += fl; d[1] += a.elapsed_time(b); d[2] += 1
rows = sorted(([k[0], k[1], v[2] // 3, v[1] / 3, v[0] / v[1] / 1e9] for k, v in agg.items()), key=lambda r: -r[3])
tot = sum(r[3] for r in rows)
print(f"{'kind':10s} {'shape':38s} {'n':>4s} {'ms/step':>8s} {'TF/s':>7s} {'us/launch':>9s}")
for r in rows:
    print(f"{r[0]:10s} {r[1]:38s} {r[2]:4d} {r[3]:8.3f} {r[4]:7.1f} {1e3 * r[3] / r[2]:9.1f}")
print("total dense ms/step", tot)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "shapes.json"), "w"))
