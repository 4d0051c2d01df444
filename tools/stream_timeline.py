"""How the two streams of the step graph overlap: times, each as its own CUDA graph, (a) the UNet encoder +
middle block (main stream before the join), (b) control net + adapter precompute (side stream), (c) the
decoder (after the join), (d) the whole dual-stream step."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
import torch
from b200sr import modules, ops
from b200sr.sampling import Stage2Engine
from oracle import configs, inputs, weights

dev = torch.device("cuda", 0)
wrapper = modules.build_stage2(configs.STAGE2_UNET, configs.STAGE2_CONTROL).eval()
weights.fill_(wrapper.state_dict(), 0)
wrapper = wrapper.to(dev)
x0, c, uc = inputs.stage2_inputs(latent=128, seed=1234)
eng = Stage2Engine(wrapper, use_graphs=True, device=dev)
eng.set_condition({k: v.to(dev) for k, v in c.items()}, {k: v.to(dev) for k, v in uc.items()})
x = x0.to(dev); noise = torch.randn_like(x)
for _ in range(3): eng.step(x, 3, noise, 0.0)
torch.cuda.synchronize()
st, cond, unet, ctl = eng._static, eng.cond, wrapper.diffusion_model, wrapper.control_model
x_hat, net_in = ops.sampler_pre(st["x"], st["noise"], st["sc"], 2)
lq = ops.nchw_to_nhwc_bf16(cond["control"].float()) if cond["control"].dtype != torch.bfloat16 else cond["control"].permute(0, 2, 3, 1).contiguous()
keep = {}
def enc():
    emb = unet._embed(st["t"], cond["vector"])
    h, hs = unet._input_stage(net_in, emb, cond["crossattn"])
    keep.update(h=unet._middle(h, emb, cond["crossattn"]), hs=hs, emb=emb)
def side():
    control = ctl.forward_nhwc(lq, st["t"], net_in, cond["crossattn"], cond["vector"])
    keep.update(control=control, pre=unet.precompute_adapters(control))
def dec():
    keep["eps"] = unet._output_stage(keep["h"], list(keep["hs"]), keep["emb"], cond["crossattn"], keep["control"], eng.control_scale,
                                     pre=keep["pre"], middle_done=True)
def timed(fn, name):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print(f"{name:45s} {t:7.3f} ms", flush=True)
    return t
res = {"encoder+middle (main, before join)": timed(enc, "encoder+middle (main, before join)"),
       "control+adapter precompute (side)": timed(side, "control+adapter precompute (side)"),
       "decoder (after join)": timed(dec, "decoder (after join)")}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.step(x, 3, noise, 0.0)
e1.record(); torch.cuda.synchronize()
res["whole step (dual stream graph)"] = e0.elapsed_time(e1) / 10
print(f"{'whole step (dual stream graph)':45s} {res['whole step (dual stream graph)']:7.3f} ms")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stream_timeline.json"), "w"), indent=1)
