"""One halo-path 3x3 convolution (2x32x32, 1280 -> 1280) and one GEGLU GEMM (2048 x 10240 x 1280) inside a
cudaProfilerStart/Stop range, for `ncu --set full --profile-from-start off`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
bf16 = torch.bfloat16
def r(*s, scale=0.5): return (torch.randn(*s, device="cuda") * scale).to(bf16)
x = r(2, 32, 32, 1280); ws = [r(1280, 9 * 1280, scale=0.02) for _ in range(3)]; b = torch.randn(1280, device="cuda"); emb = torch.randn(2, 1280, device="cuda")
a = r(2048, 1280); wg = [r(10240, 1280, scale=0.03) for _ in range(3)]; bg = torch.randn(10240, device="cuda")
for i in range(2): ops.conv3x3(x, ws[i], b, rowvec=emb); ops.gemm(a, wg[i], bg, geglu=True)
torch.cuda.synchronize(); torch.cuda.profiler.start()
ops.conv3x3(x, ws[2], b, rowvec=emb)
ops.gemm(a, wg[2], bg, geglu=True)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
