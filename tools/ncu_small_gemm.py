"""The step's most frequent launch — GEMM 2048 x 1280 x 1280 + bias + residual (293 per step) — inside a
cudaProfilerStart/Stop range, for `ncu --set full --import-source on --profile-from-start off` (source-level stalls)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
bf16 = torch.bfloat16
def r(*s, scale=0.5): return (torch.randn(*s, device="cuda") * scale).to(bf16)
a = r(2048, 1280); ws = [r(1280, 1280, scale=0.03) for _ in range(3)]; b = torch.randn(1280, device="cuda"); res = r(2048, 1280)
for i in range(2): ops.gemm(a, ws[i], b, residual=res)
torch.cuda.synchronize(); torch.cuda.profiler.start()
ops.gemm(a, ws[2], b, residual=res)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
