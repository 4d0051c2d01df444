import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
qkv = (torch.randn(2, 1024, 3840, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3): ops.attention(qkv, qkv, qkv, 20, q_col=0, k_col=1280, v_col=2560)
torch.cuda.synchronize(); torch.cuda.profiler.start()
ops.attention(qkv, qkv, qkv, 20, q_col=0, k_col=1280, v_col=2560)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
