"""In-graph time of the big GEMMs for every legal N tile (force_bn): checks pick_bn's choice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops
bf16 = torch.bfloat16
def r(*s, scale=0.5): return (torch.randn(*s, device="cuda") * scale).to(bf16)
REP = 32
def gt(fn):
    for i in range(REP): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP): fn(i)
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / REP * 1e3
for (M, N, K, geglu, res) in ((2048, 10240, 1280, True, False), (2048, 3840, 1280, False, False), (2048, 1280, 5120, False, True),
                             (2048, 1280, 1280, False, True), (8192, 5120, 640, True, False), (8192, 1920, 640, False, False), (8192, 640, 640, False, True)):
    nw = max(2, min(REP, int(160e6 / (N * K * 2)) + 1))
    ws = [r(N, K, scale=0.03) for _ in range(nw)]; b = torch.randn(N, device="cuda"); a = r(M, K)
    rs = r(M, N) if res else None
    out = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=bf16)
    line = f"M{M} N{N} K{K}{' geglu' if geglu else ''}: auto {gt(lambda i: ops.gemm(a, ws[i % nw], b, geglu=geglu, residual=rs, out=out)):.1f} us |"
    for bn in (96, 128, 160, 192, 224, 256):
        try:
            line += f" bn{bn} {gt(lambda i: ops.gemm(a, ws[i % nw], b, geglu=geglu, residual=rs, out=out, force_bn=bn)):.1f}"
        except Exception as e:
            line += f" bn{bn} -"
    print(line, flush=True)
