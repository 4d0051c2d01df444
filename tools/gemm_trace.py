"""Where a one-tile-per-CTA GEMM launch spends its time (needs a -DB200SR_GEMM_TRACE build, see tools/gpu_trace.sh).

Part 1: the MMA issue thread's mainloop (cycles per chunk, cycles blocked on data).
Part 2: the fixed cost: clock64 stamps inside every CTA of each of 8 dependent launches of the same GEMM replayed from one
CUDA graph (programmatic dependent launch edges, like the step), plus globaltimer at entry / exit to place the kernels
on one time axis: entry -> set-up done -> producer past griddepcontrol.wait -> first stage landed -> last MMA committed
-> accumulator visible -> last store -> exit, and the gap between one kernel's last exit and the next one's wait release."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
bf16 = torch.bfloat16
def r(*s, scale=0.5): return (torch.randn(*s, device="cuda") * scale).to(bf16)
def set_trace(t): lib.b200sr_debug_set_gemm_trace(ctypes.c_void_p(0 if t is None else t.data_ptr()))

trace = torch.zeros(512, 32, dtype=torch.int64, device="cuda")
for (M, N, K, geglu) in ((2048, 1280, 1280, False), (2048, 1280, 5120, False), (2048, 3840, 1280, False), (2048, 10240, 1280, True)):
    a = r(M, K); ws = [r(N, K, scale=0.03) for _ in range(6)]; b = torch.randn(N, device="cuda")
    for i in range(3): ops.gemm(a, ws[i], b, geglu=geglu)
    trace.zero_()
    set_trace(trace)
    ops.gemm(a, ws[4], b, geglu=geglu)   # cold weights
    torch.cuda.synchronize()
    set_trace(None)
    t = trace.cpu()
    t = t[t[:, 2] > 0].float()
    print(f"M{M} N{N} K{K}: CTAs issuing {len(t)}, chunks/CTA {t[:,2].mean():.0f}, mainloop {t[:,0].mean():.0f} cycles = {t[:,0].mean()/t[:,2].mean():.0f}/chunk, "
          f"blocked on data {t[:,1].mean():.0f} cycles ({100*t[:,1].mean()/t[:,0].mean():.0f}%), chunks found late {100*t[:,3].mean()/t[:,2].mean():.0f}%")

print("\nepilogue of warp 2 lane 0, cycles per 32-column chunk (tmem_ld wait | arithmetic | stores) and per tile before the accumulator wait")
def epi_case(label, fn):
    for _ in range(3): fn()
    trace.zero_(); set_trace(trace); fn(); torch.cuda.synchronize(); set_trace(None)
    t = trace.cpu(); t = t[t[:, 20] > 0].double()
    ch, tl = t[:, 20].mean(), t[:, 21].mean()
    print(f"  {label:58s} tiles/CTA {tl:5.1f} chunks/tile {ch / tl:4.1f} | ld wait {t[:,16].sum()/t[:,20].sum():6.0f} | math {t[:,17].sum()/t[:,20].sum():6.0f} | "
          f"stores {t[:,18].sum()/t[:,20].sum():6.0f} | prologue/tile {t[:,19].sum()/t[:,21].sum():6.0f}")
_a = r(2048, 1280); _w = r(1280, 1280, scale=0.03); _b = torch.randn(1280, device="cuda"); _res = r(2048, 1280)
epi_case("2048x1280x1280 bf16 + bias + residual", lambda: ops.gemm(_a, _w, _b, residual=_res))
epi_case("2048x1280x1280 bf16 + bias", lambda: ops.gemm(_a, _w, _b))
_w2 = r(10240, 1280, scale=0.03); _b2 = torch.randn(10240, device="cuda")
epi_case("2048x10240x1280 geglu", lambda: ops.gemm(_a, _w2, _b2, geglu=True))
epi_case("2048x10240x1280 bf16 + bias", lambda: ops.gemm(_a, _w2, _b2))
_q = r(16384, 512); _k = r(16384, 512); _o = torch.empty(16384, 16384, device="cuda")
epi_case("16384x16384x512 fp32 out (TMA store unless B200SR_EPI_TMA=0)", lambda: ops.gemm(_q, _k, None, alpha=0.044, out=_o, out_fp32=True))
del _o
_x = r(1, 1024, 1024, 64); _wc = r(64, 9 * 64, scale=0.02); _bc = torch.randn(64, device="cuda")
epi_case("conv3x3 1x1024x1024 64->64 + bias (halo path, resident weights)", lambda: ops.conv3x3(_x, _wc, _bc))
_x2 = r(1, 512, 512, 128); _wc2 = r(128, 9 * 128, scale=0.02); _bc2 = torch.randn(128, device="cuda")
epi_case("conv3x3 1x512x512 128->128 + bias", lambda: ops.conv3x3(_x2, _wc2, _bc2))
if os.environ.get("TRACE_EPI_ONLY"):
    raise SystemExit(0)

print("\nfixed-cost timeline, 8 dependent launches per graph (medians over CTAs, ns at the measured SM clock)")
for (M, N, K, res) in ((2048, 1280, 1280, True), (2048, 1280, 64, False)):
    NL = 8
    a = r(M, K); ws = [r(N, K, scale=0.03) for _ in range(NL)]; b = torch.randn(N, device="cuda")
    resid = r(M, N) if res else None
    outs = [torch.empty(M, N, device="cuda", dtype=bf16) for _ in range(2)]
    traces = [torch.zeros(512, 32, dtype=torch.int64, device="cuda") for _ in range(NL)]
    def body():
        for i in range(NL):
            set_trace(traces[i])
            # a dependent chain: launch i reads what launch i-1 wrote (as residual) so the wait is a real dependency
            ops.gemm(a, ws[i], b, residual=outs[(i + 1) % 2] if res else None, out=outs[i % 2])
        set_trace(None)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    ts = [t.cpu() for t in traces]
    ts = [t[t[:, 12] > 0] for t in ts]
    # SM clock from (clock64 exit - entry) / (globaltimer exit - entry) of the longest-lived CTAs
    t3 = ts[3]
    ghz = ((t3[:, 11] - t3[:, 4]).double() / (t3[:, 13] - t3[:, 12]).double().clamp(min=1)).median().item()
    print(f"\nGEMM {M}x{N}x{K}{' + residual' if res else ''}: {len(t3)} CTAs, SM clock ~{ghz:.2f} GHz")
    names = ["set-up (barriers, TMEM alloc, cluster sync)", "-> producer past griddepcontrol.wait", "-> first stage landed (MMA can start)",
             "-> last MMA committed (mainloop)", "-> accumulator visible to epilogue", "-> last store issued (epilogue)", "-> exit (syncs, TMEM dealloc)"]
    for li in (3, 4, 5):
        t = ts[li].double()
        stamps = [t[:, 4], t[:, 5], t[:, 6], t[:, 7], t[:, 8], t[:, 9], t[:, 10], t[:, 11]]
        # producer / MMA stamps exist in every CTA for slot 6; 7 and 8 only in the leader CTA of a pair
        line = []
        for k in range(7):
            a0, a1 = stamps[k], stamps[k + 1]
            ok = (a0 > 0) & (a1 > 0)
            line.append(((a1[ok] - a0[ok]).median().item() / ghz) if ok.any() else float("nan"))
        dur = (ts[li][:, 13].max() - ts[li][:, 12].min()).item()
        gap_prev = (ts[li][:, 12].min() - ts[li - 1][:, 13].max()).item()
        start_to_start = (ts[li][:, 12].min() - ts[li - 1][:, 12].min()).item()
        print(f" launch {li}: first entry to last exit {dur} ns; previous kernel's last exit -> this kernel's first entry {gap_prev} ns "
              f"(negative = launched early by PDL); start-to-start {start_to_start} ns")
        for nme, v in zip(names, line):
            print(f"    {nme:48s} {v:8.0f} ns")
        # when does the wait release relative to the previous kernel's last exit?  (globaltimer of entry + clock delta)
        rel = ts[li][:, 12].double() + (ts[li][:, 6] - ts[li][:, 4]).double() / ghz
        prev_exit = ts[li - 1][:, 13].max().item()
        ok = ts[li][:, 6] > 0
        print(f"    previous kernel's last exit -> wait released (median over CTAs) {(rel[ok] - prev_exit).median().item():8.0f} ns")
