"""Where the GEMM mainloop's issue thread spends its time (needs B200SR_EXTRA_FLAGS=-DB200SR_GEMM_TRACE ./build.sh)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
bf16 = torch.bfloat16
def r(*s, scale=0.5): return (torch.randn(*s, device="cuda") * scale).to(bf16)
trace = torch.zeros(512, 4, dtype=torch.int64, device="cuda")
for (M, N, K, geglu) in ((2048, 1280, 1280, False), (2048, 1280, 5120, False), (2048, 3840, 1280, False), (2048, 10240, 1280, True)):
    a = r(M, K); ws = [r(N, K, scale=0.03) for _ in range(6)]; b = torch.randn(N, device="cuda")
    for i in range(3): ops.gemm(a, ws[i], b, geglu=geglu)
    trace.zero_()
    lib.b200sr_debug_set_gemm_trace(ctypes.c_void_p(trace.data_ptr()))
    ops.gemm(a, ws[4], b, geglu=geglu)   # cold weights
    torch.cuda.synchronize()
    lib.b200sr_debug_set_gemm_trace(ctypes.c_void_p(0))
    t = trace.cpu()
    t = t[t[:, 2] > 0].float()
    print(f"M{M} N{N} K{K}: CTAs issuing {len(t)}, chunks/CTA {t[:,2].mean():.0f}, mainloop {t[:,0].mean():.0f} cycles = {t[:,0].mean()/t[:,2].mean():.0f}/chunk, "
          f"blocked on data {t[:,1].mean():.0f} cycles ({100*t[:,1].mean()/t[:,0].mean():.0f}%), chunks found late {100*t[:,3].mean()/t[:,2].mean():.0f}%")
