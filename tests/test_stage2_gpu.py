"""GPU parity of the stage-2 denoiser path (runs on the B200 box: pytest -m gpu).

Checker = the oracle (fp32 torch restatement, pinned against the real reference in the build
container) and the committed golden vectors produced by the real reference.
Tolerances are the ones BASELINE.json states: per-step eps rel-L2 <= 1e-2 (bf16 kernels vs fp32
reference); final latents after a sampled trajectory within PSNR >= 40 dB.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def psnr(a, b):
    a, b = a.float(), b.float()
    peak = (b.max() - b.min()).item()
    return 10 * math.log10(peak * peak / ((a - b) ** 2).mean().item())


def build(unet_cfg, control_cfg, seed=0):
    from b200sr import modules
    from oracle import weights

    w = modules.build_stage2(unet_cfg, control_cfg).eval()
    weights.fill_(w.state_dict(), seed)
    return w.cuda()


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "stage2_test_16.pt"), weights_only=False)


@pytest.fixture(scope="module")
def small():
    from oracle import configs

    return build(configs.STAGE2_UNET_TEST, configs.STAGE2_CONTROL_TEST)


def _cond(latent, dev="cuda"):
    from oracle import inputs, sampler as osampler

    x, c, uc = inputs.stage2_inputs(latent=latent, seed=1234)
    _, _, cin = osampler.cfg_prepare(x, torch.ones(1), c, uc)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}  # noqa: E731
    return x.to(dev), to(c), to(uc), to(cin)


def test_eps_matches_reference_golden(golden, small):
    """Single network call on the reference's own inputs vs the reference's fp32 output."""
    _, _, _, cin = _cond(golden["latent"])
    with torch.no_grad():
        eps = small(golden["net_x"].cuda(), golden["idx"].cuda(), cin, 1.0, "none", None)
    assert eps.dtype == torch.float32 and eps.shape == golden["eps"].shape
    err = rel_l2(eps.cpu(), golden["eps"])
    print(f"eps rel-L2 vs reference golden: {err:.4e}")
    assert err < 1e-2


def test_two_stage_protocol_and_control_scale(golden, small):
    from oracle import stage2 as ostage2

    _, _, _, cin = _cond(golden["latent"])
    x, t = golden["net_x"].cuda(), golden["idx"].cuda()
    with torch.no_grad():
        eps = small(x, t, cin, 1.0, "none", None)
        info = small(x, t, cin, 1.0, "input_stage1", None)
        assert rel_l2(info["h"].cpu(), golden["h_stage1"]) < 2e-2
        eps2 = small(x, t, cin, 1.0, "input_stage2", info)
        assert torch.equal(eps, eps2)
        sd = {k: v.detach() for k, v in small.state_dict().items()}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref = ostage2.control_wrapper(sd, x, t, cin, 0.6)
        eps_s = small(x, t, cin, 0.6, "none", None)
    assert rel_l2(eps_s, ref) < 1e-2


def test_engine_trajectory_matches_reference_golden(golden, small):
    """6 RestoreEDMSampler steps with the first-block cache: same hit/miss decisions as the real
    reference, thresholds close, final latent PSNR >= 40 dB; CUDA-graph replay == eager."""
    from b200sr.sampling import Stage2Engine

    _, c, uc, _ = _cond(golden["latent"])
    outs = []
    for graphs in (False, True):
        eng = Stage2Engine(small, use_graphs=graphs)
        assert torch.equal(eng.sched.sigmas, golden["sigmas"])
        eng.set_condition(c, uc)
        z = eng.init_latent(golden["z0"].cuda())
        thr = golden["threshold"]
        for i in range(golden["steps"]):
            torch.manual_seed(1000 + i)
            noise = torch.randn(golden["z0"].shape).cuda()  # same CPU stream the reference drew from
            z, thr = eng.step(z, i, noise, thr)
        outs.append(z)
        print("trace", eng.trace, "launches", eng.launches)
        assert [t[0] for t in eng.trace] == [t[0] for t in golden["trace"]]
        assert [t[1] for t in eng.trace] == pytest.approx([t[1] for t in golden["trace"]], rel=5e-2)
        assert psnr(z.cpu(), golden["z_final"]) >= 40.0
    assert torch.equal(outs[0], outs[1])


def test_uncached_step_engine_vs_oracle(small):
    """threshold <= 0 path (the bench workload) at a non-trivial size, oracle evaluated on the GPU in fp32."""
    from b200sr.sampling import Stage2Engine
    from oracle import sampler as osampler, stage2 as ostage2

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    latent = 32
    x, c, uc, _ = _cond(latent)
    sd = {k: v.detach() for k, v in small.state_dict().items()}
    eng = Stage2Engine(small)
    eng.set_condition(c, uc)
    oden = osampler.Denoiser(device="cuda")
    net = lambda xx, tt, cc, cs, mode, pi: ostage2.control_wrapper(sd, xx, tt, cc, cs)  # noqa: E731
    smp = osampler.RestoreSampler(device="cuda")
    _, s_in, sig = smp.init_loop(x.clone())
    g = torch.Generator(device="cuda").manual_seed(3)
    noise = torch.randn(x.shape, generator=g, device="cuda")
    with torch.no_grad():
        ref, _ = smp.step(x, 3, s_in, sig, lambda *a: oden(net, *a), c, uc, 1.0, 0.0, None, noise)
    out, _ = eng.step(x, 3, noise, 0.0)
    # compare the update direction (x_next - x), which is what the network contributes
    assert rel_l2(out - x, ref - x) < 1e-2
    assert psnr(out, ref) >= 40.0
    # the other engine schedules compute the same step within the same tolerance: single stream without graphs (the
    # adapters' control-side work is then fused differently, so not bit-identical), CFG halves on separate streams
    for kw in ({"dual_stream": False, "use_graphs": False}, {"split_cfg": True}):
        alt = Stage2Engine(small, **kw)
        alt.set_condition(c, uc)
        out_alt, _ = alt.step(x, 3, noise, 0.0)
        assert rel_l2(out_alt - x, ref - x) < 1e-2
        assert rel_l2(out_alt - x, out - x) < 5e-3


@pytest.fixture(scope="module")
def full():
    from oracle import configs

    return build(configs.STAGE2_UNET, configs.STAGE2_CONTROL)


def test_full_model_eps_1024(full):
    """BASELINE config 2: full SDXL UNet + ControlNet, 128^2 latent, CFG batch 2, vs the fp32 oracle."""
    from oracle import sampler as osampler, stage2 as ostage2

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    x, c, uc, cin = _cond(128)
    sig = torch.full((1,), 14.6146, device="cuda")
    oden = osampler.Denoiser(device="cuda")
    xin = torch.cat([x] * 2)
    idx = oden.sigma_to_idx(torch.cat([sig] * 2))
    net_x = xin / (oden.sigmas[idx] ** 2 + 1).sqrt().view(-1, 1, 1, 1)
    sd = {k: v.detach() for k, v in full.state_dict().items()}
    with torch.no_grad():
        ref = ostage2.control_wrapper(sd, net_x, idx, cin, 1.0)
        eps = full(net_x, idx, cin, 1.0, "none", None)
    err = rel_l2(eps, ref)
    print(f"full-model eps rel-L2 vs fp32 oracle: {err:.4e}; |eps| std {ref.std().item():.3f}")
    assert err < 1e-2
    # The engine binds the text context once per image (K/V hoisted, to_q / to_out folded into them: every
    # text cross-attention becomes two GEMMs).  Same tolerance against the same oracle output.
    from b200sr.modules import CrossAttention, bind_text_context

    cin_b = dict(cin)
    cin_b["crossattn"] = ctx = cin["crossattn"].to(torch.bfloat16).contiguous()  # bound by tensor identity
    bind_text_context(full, ctx)
    try:
        folded = [m for m in full.modules() if isinstance(m, CrossAttention) and m.__dict__.get("_static")]
        assert len(folded) >= 90 and all(v[2] is not None for m in folded for v in m.__dict__["_static"].values())
        with torch.no_grad():
            eps_b = full(net_x, idx, cin_b, 1.0, "none", None)
    finally:
        bind_text_context(full, None)
    err_b = rel_l2(eps_b, ref)
    print(f"full-model eps rel-L2 vs fp32 oracle, text context bound (folded cross-attention): {err_b:.4e}")
    assert err_b < 1e-2


def _oracle_on_gpu(model):
    from oracle import sampler as osampler, stage2 as ostage2

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    oden = osampler.Denoiser(device="cuda")
    net = ostage2.network(sd)
    return (lambda *a: oden(net, *a)), osampler.RestoreSampler(device="cuda")


def test_config3_full_trajectory_psnr(full):
    """BASELINE config 3, stage 2: 50 RestoreEDMSampler steps of the FULL-depth model on a 128^2 latent with the
    first-block cache (img_threshold 0.3, dec 1), CUDA graphs, vs the fp32 oracle run on the same GPU with the
    same noises.  Gate: identical hit/miss trace and final latent PSNR >= 40 dB.
    A trace that diverges at a near-tie of the similarity test (|diff - thr| within the bf16 error of diff) is
    reported as such — xfail with the step and margin — not silently passed; any other divergence fails."""
    from b200sr.sampling import Stage2Engine
    from oracle import sampler as osampler

    x, c, uc, _ = _cond(128)
    steps, threshold = 50, 0.3
    g = torch.Generator(device="cuda").manual_seed(2024)
    z0 = torch.randn(1, 4, 128, 128, generator=g, device="cuda")
    noises = [torch.randn(1, 4, 128, 128, generator=g, device="cuda") for _ in range(steps)]
    den, smp = _oracle_on_gpu(full)
    zo, s_in, sig = smp.init_loop(z0.clone())
    thr, cache = threshold, osampler.CacheState()
    with torch.no_grad():
        for i in range(steps):
            zo, thr = smp.step(zo, i, s_in, sig, den, c, uc, 1.0, thr, cache, noises[i])
    eng = Stage2Engine(full)
    eng.set_condition(c, uc)
    z = eng.sample(z0, noises, threshold=threshold, dec=1.0)
    ours, ref = [t[0] for t in eng.trace], [t[0] for t in smp.trace]
    print("oracle trace:", "".join("H" if t == "hit" else "m" for t in ref), " misses", ref.count("miss"))
    print("b200sr trace:", "".join("H" if t == "hit" else "m" for t in ours), " misses", ours.count("miss"))
    if ours != ref:
        k = next(i for i, (a, b) in enumerate(zip(ours, ref)) if a != b)
        d_o, d_r = eng.trace[k][1], smp.trace[k][1]
        thr_k = next((t[1] for t in reversed(smp.trace[:k]) if t[0] == "miss"), threshold)
        margin = abs(d_r - thr_k) / max(thr_k, 1e-12)
        msg = (f"hit/miss traces diverge at step {k}: oracle diff {d_r:.6f}, b200sr diff {d_o:.6f}, threshold {thr_k:.6f} "
               f"(relative margin {margin:.3e})")
        assert margin < 2e-2, msg + " — not a near-tie"
        pytest.xfail(msg + " — near-tie of the similarity test; trajectories are not comparable beyond it")
    assert [t[1] for t in eng.trace] == pytest.approx([t[1] for t in smp.trace], rel=1e-1)
    p = psnr(z, zo)
    print(f"50-step latent PSNR vs fp32 oracle: {p:.2f} dB, rel-L2 {rel_l2(z, zo):.3e}")
    assert p >= 40.0


def test_tiled_step_vs_oracle(full):
    """BASELINE config 4 on one GPU: one tiled sampler step (sampling.py:716-756 with SURVEY section 5 semantics) on
    a 256^2 latent (9 windows of 128^2, stride 96) vs the fp32 oracle's tiled step; also with three windows per
    network call (shared caption), and through the rank-sharded stepper (world 1)."""
    from b200sr import ops
    from b200sr.parallel import EngineTileRunner, PooledTileStepper
    from b200sr.sampling import Stage2Engine
    from oracle import sampler as osampler

    L = 256
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = torch.randn(1, 4, L, L, generator=g, device="cuda") * (1 + 14.6146**2) ** 0.5
    lq = torch.randn(1, 4, L, L, generator=g, device="cuda")
    noise = torch.randn(1, 4, L, L, generator=g, device="cuda")
    _, c, uc, _ = _cond(128)
    c = {k: v for k, v in c.items() if k != "control"}
    uc = {k: v for k, v in uc.items() if k != "control"}
    den, smp = _oracle_on_gpu(full)
    _, s_in, sig = smp.init_loop(x.clone())
    i = 3
    with torch.no_grad():
        ref = osampler.tiled_step(smp, x, i, s_in, sig, den, dict(c, control=lq), dict(uc, control=lq), 128, 96, noise)
    eng = Stage2Engine(full)
    out1 = eng.tiled_step(x, i, noise, lq, c, uc, 128, 96)
    print(f"tiled step rel-L2 of the update vs oracle: {rel_l2(out1 - x, ref - x):.3e}, PSNR {psnr(out1, ref):.1f} dB")
    assert rel_l2(out1 - x, ref - x) < 1e-2 and psnr(out1, ref) >= 40.0
    out3 = Stage2Engine(full).tiled_step(x, i, noise, lq, c, uc, 128, 96, tile_batch=3)
    assert rel_l2(out3 - x, ref - x) < 1e-2 and rel_l2(out3 - x, out1 - x) < 5e-3
    # the sharded stepper with world size 1 is the same computation: bit-identical to tiled_step
    st = PooledTileStepper(1, L, L, 128, 96, device="cuda", blend=ops)
    runner = EngineTileRunner(lambda: Stage2Engine(full), {0: (c, uc)}, {0: lq})
    out_s = st.step({0: x}, i, {0: noise}, runner)[0]
    assert torch.equal(out_s, out1)


def test_engine_batch_of_latents(small):
    """Two latents per step (CFG batch 4) == the latents run alone, within bf16 noise; the step loader's
    precomputed embedding rows == the per-step embedding path."""
    from b200sr.sampling import Stage2Engine
    from oracle import inputs

    xa, ca, uca = inputs.stage2_inputs(latent=32, seed=21)
    xb, cb, ucb = inputs.stage2_inputs(latent=32, seed=22)
    to = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    ca, uca, cb, ucb, xa, xb = to(ca), to(uca), to(cb), to(ucb), xa.cuda(), xb.cuda()
    g = torch.Generator(device="cuda").manual_seed(9)
    na, nb = torch.randn(xa.shape, generator=g, device="cuda"), torch.randn(xa.shape, generator=g, device="cuda")
    e1 = Stage2Engine(small)
    e1.set_condition(ca, uca)
    ra, _ = e1.step(xa, 5, na, 0.0)
    e1.set_condition(cb, ucb)
    rb, _ = e1.step(xb, 5, nb, 0.0)
    e2 = Stage2Engine(small)
    cat = lambda a, b: {k: torch.cat((a[k], b[k]), 0) for k in a}  # noqa: E731
    e2.set_condition(cat(ca, cb), cat(uca, ucb))
    rab, _ = e2.step(torch.cat((xa, xb), 0), 5, torch.cat((na, nb), 0), 0.0)
    assert rel_l2(rab[:1] - xa, ra - xa) < 5e-3 and rel_l2(rab[1:] - xb, rb - xb) < 5e-3
    e3 = Stage2Engine(small, precompute_emb=False)
    e3.set_condition(ca, uca)
    rc, _ = e3.step(xa, 5, na, 0.0)
    assert rel_l2(rc - xa, ra - xa) < 2e-3


def test_lazy_control_matches_eager_control(golden, small):
    """First-block cache with the control net deferred to the miss path (a hit then runs the UNet encoder only) vs
    computing it in the first call like the reference (wrappers.py:91-95): same hit/miss decisions, similarity values
    within bf16 noise (the miss path folds the adapters' control-side work differently in the two schedules, so the bits
    differ), final latents within PSNR >= 40 dB of each other and of the reference's golden trajectory."""
    from b200sr.sampling import Stage2Engine

    _, c, uc, _ = _cond(golden["latent"])
    outs, traces, launches = [], [], []
    for lazy in (True, False):
        eng = Stage2Engine(small, lazy_control=lazy)
        eng.set_condition(c, uc)
        z = eng.init_latent(golden["z0"].cuda())
        thr = golden["threshold"]
        for i in range(golden["steps"]):
            torch.manual_seed(1000 + i)
            noise = torch.randn(golden["z0"].shape).cuda()
            z, thr = eng.step(z, i, noise, thr)
        outs.append(z)
        traces.append(eng.trace)
        launches.append(dict(eng.launches))
        eng.close()
    print("launches lazy / eager control:", launches)
    assert [t[0] for t in traces[0]] == [t[0] for t in traces[1]] == [t[0] for t in golden["trace"]]
    assert [t[1] for t in traces[0]] == pytest.approx([t[1] for t in traces[1]], rel=5e-3)
    assert psnr(outs[0], outs[1]) >= 40.0
    assert psnr(outs[0].cpu(), golden["z_final"]) >= 40.0 and psnr(outs[1].cpu(), golden["z_final"]) >= 40.0
    assert launches[0]["stage1"] < launches[1]["stage1"]   # a hit no longer pays for the control net
