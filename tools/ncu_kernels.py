"""Launch a handful of representative dense kernels once each (after warm-up) for `ncu --set full`."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
bf16 = torch.bfloat16
dev = "cuda"
def r(*s): return (torch.randn(*s, device=dev) * 0.5).to(bf16)
cases = []
a = r(2048, 1280); w = r(1280, 1280); b = torch.randn(1280, device=dev); res = r(2048, 1280)
cases.append(("gemm_2048x1280x1280_res", lambda: ops.gemm(a, w, b, residual=res)))
wg, bg = ops.pack_geglu(torch.randn(10240, 1280, device=dev) * 0.03, torch.randn(10240, device=dev) * 0.1)
cases.append(("gemm_geglu_2048x10240x1280", lambda: ops.gemm(a, wg, bg, geglu=True)))
a5 = r(2048, 5120); w5 = r(1280, 5120)
cases.append(("gemm_2048x1280x5120_res", lambda: ops.gemm(a5, w5, b, residual=res)))
wq = r(3840, 1280)
cases.append(("gemm_qkv_2048x3840x1280", lambda: ops.gemm(a, wq)))
qkv = r(2, 1024, 3840)
cases.append(("attn_1024", lambda: ops.attention(qkv, qkv, qkv, 20, q_col=0, k_col=1280, v_col=2560)))
qkv4 = r(2, 4096, 1920)
cases.append(("attn_4096", lambda: ops.attention(qkv4, qkv4, qkv4, 10, q_col=0, k_col=640, v_col=1280)))
x = r(2, 32, 32, 1280); wc = r(1280, 9 * 1280); emb = torch.randn(2, 1280, device=dev)
cases.append(("conv_32x32_1280", lambda: ops.conv3x3(x, wc, b, rowvec=emb)))
xg = r(2, 128, 128, 320); g = torch.ones(320, device=dev); z = torch.zeros(320, device=dev)
cases.append(("gn_silu_128x128x320", lambda: ops.group_norm(xg, g, z, silu=True)))
for name, fn in cases:
    for _ in range(3): fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in cases:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", [c[0] for c in cases])
