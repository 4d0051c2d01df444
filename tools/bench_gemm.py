"""GEMM micro-benchmark: b200sr tcgen05 kernel vs cuBLAS (torch.matmul) on the step's shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
bf16 = torch.bfloat16
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us
shapes = [(8192, 8192, 8192), (2048, 10240, 1280), (2048, 1280, 5120), (2048, 3840, 1280), (2048, 1280, 1280),
          (8192, 5120, 640), (8192, 640, 640), (8192, 640, 2560), (32768, 320, 320)]
bns = [int(a) for a in sys.argv[1:]] or [0]
print(f"{'M':>6s} {'N':>6s} {'K':>6s} | {'cuBLAS us':>9s} {'TF/s':>7s} | " + " | ".join(f"bn={b:<3d} us   TF/s" for b in bns))
for M, N, K in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(bf16); w = (torch.randn(N, K, device="cuda") * 0.05).to(bf16)
    out = torch.empty(M, N, device="cuda", dtype=bf16)
    fl = 2.0 * M * N * K
    t_ref = timeit(lambda: torch.matmul(a, w.t(), out=out))
    cols = []
    for bn in bns:
        if bn > N: cols.append("      -       -"); continue
        t = timeit(lambda: ops.gemm(a, w, None, out=out, force_bn=bn))
        cols.append(f"{t:9.1f} {fl / t / 1e6:7.1f}")
    print(f"{M:6d} {N:6d} {K:6d} | {t_ref:9.1f} {fl / t_ref / 1e6:7.1f} | " + " | ".join(cols))
