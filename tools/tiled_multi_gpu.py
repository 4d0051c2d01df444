"""Config 4 on real GPUs: one 256^2 latent (2048^2 image), 128^2 tiles with stride 96 (9 windows) sharded over
the ranks with NVLink P2P halo exchange (NCCL send/recv).  Checks N-GPU == 1-GPU and reports steps/s.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tiled_multi_gpu.py [--steps K] [--test-config]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
import torch.distributed as dist
from b200sr import modules, ops
from b200sr.parallel import TileShardedStepper
from b200sr.sampling import Stage2Engine
from oracle import configs, inputs, weights

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--latent", type=int, default=256)
ap.add_argument("--test-config", action="store_true", help="reduced transformer depth (fast build)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ucfg, ccfg = (configs.STAGE2_UNET_TEST, configs.STAGE2_CONTROL_TEST) if args.test_config else (configs.STAGE2_UNET, configs.STAGE2_CONTROL)
w = modules.build_stage2(ucfg, ccfg).eval()
weights.fill_(w.state_dict(), 0)
w = w.to(dev)
L = args.latent
g = torch.Generator().manual_seed(4321)                     # identical on every rank
x = torch.randn(1, 4, L, L, generator=g) * (1 + 14.6146**2) ** 0.5
lq = torch.randn(1, 4, L, L, generator=g)
c = {"crossattn": torch.randn(1, 77, 2048, generator=g), "vector": torch.randn(1, 2816, generator=g)}
uc = {"crossattn": torch.randn(1, 77, 2048, generator=g), "vector": torch.randn(1, 2816, generator=g)}
noises = [torch.randn(1, 4, L, L, generator=g) for _ in range(args.steps)]
x, lq = x.to(dev), lq.to(dev)
c = {k: v.to(dev) for k, v in c.items()}; uc = {k: v.to(dev) for k, v in uc.items()}
noises = [n.to(dev) for n in noises]
eng = Stage2Engine(w, device=dev)
_cnt_scratch = torch.zeros(1, 4, L, L, device=dev)


def accumulate(tile, weight, acc, h0, w0):
    ops.tile_accumulate(tile, weight, acc, _cnt_scratch, h0, w0)   # count is data independent: kept by the stepper


def run(stepper, sync=True):
    xx = x.clone()
    def step_fn(x_tile, i, noise_tile, win):
        h0, h1, w0, w1 = win
        lqt = lq[:, :, h0:h1, w0:w1].contiguous()
        eng.set_condition(dict(c, control=lqt), dict(uc, control=lqt))
        out, _ = eng.step(x_tile, i, noise_tile, 0.0)
        return out
    torch.cuda.synchronize()
    if world > 1 and sync: dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        xx = stepper.step(xx, i, noises[i], step_fn, accumulate)
    torch.cuda.synchronize()
    if world > 1 and sync: dist.barrier()
    return xx, (time.perf_counter() - t0) / args.steps

st = TileShardedStepper(L, L, 128, 96, device=dev)
run(st)                                                     # warm-up (graph capture, weight packing)
out, dt = run(st)
full = st.gather_full(out)
res = {"n_gpus": world, "windows": len(st.windows), "windows_per_rank": [len(p) for p in st.parts],
       "s_per_tiled_step": dt, "tiled_steps_per_s": 1.0 / dt, "halo_bytes_per_step_rank0": st.halo_bytes_per_step}
if world > 1:
    # reference: rank 0 recomputes all windows alone and compares
    solo = TileShardedStepper.__new__(TileShardedStepper)
    solo.__dict__.update(st.__dict__)
    solo.world, solo.rank, solo.mine, solo.plan = 1, 0, list(st.windows), {}
    if rank == 0:
        ref, dt1 = run(solo, sync=False)
        res["max_abs_diff_vs_1gpu"] = (full - ref).abs().max().item()
        res["rel_l2_vs_1gpu"] = ((full - ref).norm() / ref.norm()).item()
        res["s_per_tiled_step_1gpu"] = dt1
        res["speedup"] = dt1 / dt
    dist.barrier()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
