"""Timeline of one attention CTA from clock64 stamps (needs a library built with
B200SR_EXTRA_FLAGS=-DB200SR_ATT_TRACE ./build.sh; debug only, never a bench number)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = 20 if N == 1024 else 10
qkv = (torch.randn(2, N, 3 * H * 64, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3): ops.attention(qkv, qkv, qkv, H, q_col=0, k_col=H * 64, v_col=2 * H * 64)
trace_all = torch.zeros(20480 + 3 * 4096, dtype=torch.int64, device="cuda")
trace = trace_all[:20480].view(4, 10, 64, 8)
lib.b200sr_debug_set_attn_trace(ctypes.c_void_p(trace_all.data_ptr()))
ops.attention(qkv, qkv, qkv, H, q_col=0, k_col=H * 64, v_col=2 * H * 64)
torch.cuda.synchronize()
lib.b200sr_debug_set_attn_trace(ctypes.c_void_p(0))
t = trace.cpu()
nblk = N // 128
for cta in range(2):
    base = int(t[cta][t[cta] > 0].min())
    rel = lambda v: int(v) - base if v > 0 else -1
    print(f"--- CTA {cta}: cycles since its first stamp; softmax warp 2 (half 0) and warp 6 (half 1)")
    print("blk | tma:kv_empty | mma: p_full  pv_issued  k_full  qk_issued | w2: s_full ld_done max_xch exps_done st_done arrived | w6: s_full ... arrived")
    for j in range(min(nblk, 10)):
        tm = rel(t[cta, 0, j, 0])
        mm = [rel(t[cta, 1, j, e]) for e in range(4)]
        w2 = [rel(t[cta, 2, j, e]) for e in range(6)]
        w6 = [rel(t[cta, 6, j, e]) for e in range(6)]
        print(f"{j:3d} | {tm:6d} | {mm} | {w2} | {w6}")
    # per-block period and phase durations averaged over the warps and blocks 1..nblk-1
    import statistics
    per, wait_s, ld, xch, exps, st = [], [], [], [], [], []
    for w in range(2, 10):
        for j in range(1, nblk):
            a = [int(t[cta, w, j, e]) for e in range(6)]; prev = int(t[cta, w, j - 1, 5])
            per.append(a[5] - prev); wait_s.append(a[0] - prev); ld.append(a[1] - a[0]); xch.append(a[2] - a[1]); exps.append(a[3] - a[2]); st.append(a[5] - a[3])
    m = statistics.mean
    print(f"mean per block: period {m(per):.0f}  wait s_full {m(wait_s):.0f}  tmem ld {m(ld):.0f}  max+exchange {m(xch):.0f}  exps {m(exps):.0f}  P store+arrive {m(st):.0f}")
    pv = [int(t[cta, 1, j, 1]) - int(t[cta, 1, j, 0]) for j in range(nblk)]
    lat = [int(t[cta, 2, j + 1, 0]) - int(t[cta, 1, j, 0]) for j in range(nblk - 1)]
    print(f"mma warp: p_full->PV issued {m(pv):.0f}; p_full seen by MMA -> next s_full seen by softmax {m(lat):.0f}")
    arr = [int(t[cta, 1, j, 0]) - max(int(t[cta, w, j, 5]) for w in range(2, 10)) for j in range(nblk)]
    print(f"last softmax arrive -> MMA warp sees p_full: {m(arr):.0f}")

# whole-CTA schedule (globaltimer ns): rounds and per-SM occupancy
ncta = 2 * H * (N // 128)
c = trace_all[20480:20480 + 3 * ncta].view(ncta, 3).cpu()
t0 = int(c[:, 0].min())
start = (c[:, 0] - t0).float() / 1e3; end = (c[:, 1] - t0).float() / 1e3; dur = end - start
print(f"CTAs {ncta}: kernel span {float(end.max()):.1f} us; CTA duration mean {float(dur.mean()):.1f} min {float(dur.min()):.1f} max {float(dur.max()):.1f} us")
order = torch.argsort(start)
for name, sel in (("first 296 started", order[:296]), ("rest", order[296:])):
    if len(sel): print(f"  {name}: start {float(start[sel].min()):.1f}..{float(start[sel].max()):.1f} us, duration mean {float(dur[sel].mean()):.1f} us, end max {float(end[sel].max()):.1f}")
sm_counts = torch.bincount(c[:, 2].int())
print("  CTAs per SM: min", int(sm_counts.min()), "max", int(sm_counts.max()), "SMs used", int((sm_counts > 0).sum()))
