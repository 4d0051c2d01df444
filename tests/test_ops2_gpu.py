"""GPU parity of the round-2 operator additions against plain torch fp32 references: row softmax (warp-per-row and the
single-pass CTA-per-row kernel for long rows), the chunked single-head attention helper, copy_batch, weighted strips."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("rows,cols,valid", [(37, 80, 77), (64, 1024, 1024), (33, 4096, 4096), (16, 16384, 16384),
                                             (8, 16384, 16001), (4, 32768, 32768), (3, 65536, 65536), (5, 1028, 1000)])
def test_softmax_rows(rows, cols, valid):
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(rows * 7 + cols)
    x = torch.randn(rows, cols, generator=g, device="cuda") * 6
    y = ops.softmax_rows(x, 0.37, valid_cols=valid)
    ref = torch.zeros_like(x)
    ref[:, :valid] = torch.softmax(x[:, :valid] * 0.37, dim=-1)
    assert y.dtype == bf16 and torch.equal(y[:, valid:], torch.zeros_like(y[:, valid:]))
    assert rel_l2(y, ref) < 4e-3
    assert torch.allclose(y.float().sum(-1), torch.ones(rows, device="cuda"), atol=2e-2)


@pytest.mark.parametrize("t,chunk", [(1024, 1 << 30), (4096, 4 * 4096 * 1024)])
def test_single_head_attention_chunked(t, chunk, monkeypatch):
    from b200sr import ops

    monkeypatch.setattr(ops, "SCORE_CHUNK_BYTES", chunk)      # second case: 4 query chunks
    g = torch.Generator(device="cuda").manual_seed(t)
    c = 512
    q = (torch.randn(t, c, generator=g, device="cuda") * 0.5).to(bf16)
    k = (torch.randn(t, c, generator=g, device="cuda") * 0.5).to(bf16)
    v = (torch.randn(t, c, generator=g, device="cuda") * 0.5).to(bf16)
    bias = torch.randn(c, generator=g, device="cuda")
    o = ops.single_head_attention(q, k, v.t().contiguous(), c ** -0.5, out_bias=bias)
    ref = torch.softmax(q.float() @ k.float().t() * c ** -0.5, -1) @ v.float() + bias
    assert rel_l2(o, ref) < 1e-2


def test_copy_batch_and_strips():
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    srcs = [torch.randn(n, generator=g, device="cuda") for n in (6, 1000, 4096, 3, 65536, 7, 128, 12)]
    dsts = [torch.zeros_like(s) for s in srcs]
    ops.copy_batch(list(zip(dsts, srcs)))
    assert all(torch.equal(d, s) for d, s in zip(dsts, srcs))
    h = torch.randn(2, 8, 16, generator=g, device="cuda").to(bf16)
    d = torch.zeros_like(h)
    ops.copy_batch([(d, h)])
    assert torch.equal(d, h)
    tile = torch.randn(1, 4, 32, 32, generator=g, device="cuda")
    w = torch.rand(32, 32, generator=g, device="cuda")
    acc1, acc2 = torch.zeros(1, 4, 64, 64, device="cuda"), torch.zeros(1, 4, 64, 64, device="cuda")
    ops.tile_accumulate(tile, w, acc1, None, 8, 24)
    # the same contribution shipped as two strips and added on the "owner": identical bits
    for (y0, x0, sh, sw) in ((0, 0, 32, 10), (0, 10, 32, 22)):
        ops.strip_add(ops.tile_weighted_strip(tile, w, y0, x0, sh, sw), acc2, 8 + y0, 24 + x0)
    assert torch.equal(acc1, acc2)
    assert torch.equal(acc1[:, :, 8:40, 24:56], tile * w)


@pytest.mark.parametrize("n,h,w,cin,cout,extras", [
    (2, 32, 32, 1280, 1280, "rowvec"), (2, 128, 128, 320, 320, "residual"), (1, 64, 64, 640, 320, ""),
    (1, 16, 8, 64, 64, ""), (1, 256, 256, 64, 64, "residual"), (3, 48, 40, 128, 256, "rowvec"), (1, 1024, 64, 128, 64, ""),
])
def test_conv3x3_with_fused_input_groupnorm(n, h, w, cin, cout, extras):
    """GroupNorm + SiLU applied to the staged input tiles inside the halo convolution == GroupNorm kernel followed by the
    convolution, bit for bit (same arithmetic, same bf16 rounding point), and == a torch fp32 reference within bf16."""
    import torch.nn.functional as F
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h + cin)
    x = (torch.randn(n, h, w, cin, generator=g, device="cuda") * 1.5 + 0.3).to(bf16)
    wt = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") * (1.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    gw = 1 + 0.1 * torch.randn(cin, generator=g, device="cuda")
    gb = 0.1 * torch.randn(cin, generator=g, device="cuda")
    kw = {}
    if extras == "rowvec":
        kw["rowvec"] = torch.randn(n, cout, generator=g, device="cuda")
    if extras == "residual":
        kw["residual"] = torch.randn(n, h, w, cout, generator=g, device="cuda").to(bf16)
    assert ops.conv3x3_gn_fusable(x)
    wp = ops.pack_conv3x3(wt)
    two = ops.conv3x3(ops.group_norm(x, gw, gb, groups=32, eps=1e-5, silu=True), wp, b, **kw)
    stats = ops.group_norm_stats(x, 32, 1e-5)
    fused = ops.conv3x3(x, wp, b, gn=(stats, gw, gb, 32, True), **kw)
    assert torch.equal(fused, two)
    xf = x.float().permute(0, 3, 1, 2)
    y = F.silu(F.group_norm(xf, 32, gw, gb, 1e-5)).to(bf16).float()
    ref = F.conv2d(y, wt.to(bf16).float(), b, padding=1).permute(0, 2, 3, 1)
    if "rowvec" in kw:
        ref = ref + kw["rowvec"][:, None, None, :]
    if "residual" in kw:
        ref = ref + kw["residual"].float()
    assert rel_l2(fused, ref) < 5e-3
    mean = xf.reshape(n, 32, -1).mean(-1)
    assert torch.allclose(stats[..., 0], mean, atol=2e-3)
