"""One eager stage-2 step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off),
plus the ordered list of C-ABI calls (one line per kernel) to join with ncu's launch list."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import _lib, modules, ops
from b200sr.sampling import Stage2Engine
from oracle import configs, inputs

dev = torch.device("cuda")
w = modules.build_stage2(configs.STAGE2_UNET, configs.STAGE2_CONTROL).eval().to(dev)  # default init is fine for timing
x0, c, uc = inputs.stage2_inputs(latent=128, seed=1234)
eng = Stage2Engine(w, use_graphs=False, device=dev)
eng.set_condition({k: v.to(dev) for k, v in c.items()}, {k: v.to(dev) for k, v in uc.items()})
x = x0.to(dev); noise = torch.randn_like(x)
for _ in range(2):
    eng.step(x, 3, noise, 0.0)
torch.cuda.synchronize()
_lib.TRACE = []
torch.cuda.profiler.start()
eng.step(x, 3, noise, 0.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
json.dump(_lib.TRACE, open(os.path.join(ROOT, "gpurun_out", "step_calls.json"), "w"))
print("kernels in step:", len(_lib.TRACE))
