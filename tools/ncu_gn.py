import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
def r(*s): return (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
cases = [(2, 1024, 1280), (2, 16384, 320), (2, 4096, 640)]
xs = [(r(n, hw, c), torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")) for n, hw, c in cases]
x2 = r(2048, 1280); g2 = torch.ones(1280, device="cuda"); b2 = torch.zeros(1280, device="cuda")
for _ in range(3):
    for x, g, b in xs: ops.group_norm(x, g, b, silu=True)
    ops.layer_norm(x2, g2, b2)
torch.cuda.synchronize(); torch.cuda.profiler.start()
for x, g, b in xs: ops.group_norm(x, g, b, silu=True)
ops.layer_norm(x2, g2, b2)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
