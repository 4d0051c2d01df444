"""GPU parity of the conditioner's text towers (causal attention, GELU epilogues, embedding gather)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
bf16 = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_causal_attention_and_gelu_epilogues():
    from b200sr import ops

    g = torch.Generator(device="cuda").manual_seed(4)
    for (b, h, t) in ((2, 12, 77), (1, 20, 77), (2, 4, 128), (1, 2, 16)):
        c = h * 64
        qkv = (torch.randn(b, t, 3 * c, generator=g, device="cuda") * 0.7).to(bf16)
        o = ops.attention(qkv, qkv, qkv, h, q_col=0, k_col=c, v_col=2 * c, scale=0.125, causal=True)
        q, k, v = (z.float().reshape(b, t, h, 64).transpose(1, 2) for z in qkv.chunk(3, dim=-1))
        s = q @ k.transpose(-1, -2) * 0.125 + torch.full((t, t), float("-inf"), device="cuda").triu(1)
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b, t, c)
        assert rel_l2(o, ref) < 1e-2, (b, h, t)
    a = (torch.randn(154, 768, generator=g, device="cuda") * 0.5).to(bf16)
    w = (torch.randn(3072, 768, generator=g, device="cuda") * 0.05).to(bf16)
    bias = torch.randn(3072, generator=g, device="cuda") * 0.1
    y = a.float() @ w.float().t() + bias
    assert rel_l2(ops.gemm(a, w, bias, act=2), torch.nn.functional.gelu(y)) < 5e-3
    assert rel_l2(ops.gemm(a, w, bias, act=3), y * torch.sigmoid(1.702 * y)) < 5e-3
    ids = torch.randint(0, 1000, (3, 77), generator=g, device="cuda")
    tok, pos = torch.randn(1000, 768, generator=g, device="cuda"), torch.randn(77, 768, generator=g, device="cuda")
    assert torch.equal(ops.embed_tokens(ids, tok, pos), (tok[ids] + pos[None]).to(bf16))


def test_conditioner_matches_transformers_golden():
    from b200sr import conditioner
    from oracle import weights

    g = torch.load(os.path.join(GOLDEN, "conditioner_small.pt"), weights_only=False)
    m = conditioner.GeneralConditionerWithControl(_clip_layers=g["clip_layers"], _clip_layer_idx=g["clip_idx"],
                                                  _bigg_layers=g["bigg_layers"]).eval()
    weights.fill_(m.state_dict(), 0)
    m = m.cuda()
    ids = g["ids"].cuda()
    batch = {"txt": (ids, ids.flip(0)), "original_size_as_tuple": torch.tensor([[1024., 1024.]] * 2, device="cuda"),
             "crop_coords_top_left": torch.zeros(2, 2, device="cuda"),
             "target_size_as_tuple": torch.tensor([[1024., 1024.]] * 2, device="cuda")}
    out = m(batch)
    e1, e2 = rel_l2(out["crossattn"].cpu(), g["crossattn"]), rel_l2(out["vector"].cpu(), g["vector"])
    print(f"conditioner (reduced depth) crossattn rel-L2 {e1:.3e}, vector rel-L2 {e2:.3e}")
    assert e1 < 1e-2 and e2 < 1e-2


def test_full_depth_conditioner_vs_fp32_oracle():
    """The shipped depths (CLIP-L 12 layers / hidden 11, bigG 32 layers) vs the fp32 oracle on the same GPU."""
    from b200sr import conditioner
    from oracle import conditioner as ocond, weights

    torch.backends.cuda.matmul.allow_tf32 = False
    m = conditioner.GeneralConditionerWithControl().eval()
    weights.fill_(m.state_dict(), 0)
    m = m.cuda()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(8)
    ids = torch.randint(1, 49000, (2, 77), generator=gen)
    ids[0, 40], ids[1, 76] = 49407, 49407
    ids[0, 41:] = 0
    ids = ids.cuda()
    z = torch.zeros(2, 4, 128, 128, device="cuda")
    c, uc = conditioner.prepare_condition(m, z, (ids, ids), (ids.flip(0), ids.flip(0)))
    batch = {"txt": (ids, ids), "original_size_as_tuple": torch.tensor([[1024., 1024.]] * 2, device="cuda"),
             "crop_coords_top_left": torch.zeros(2, 2, device="cuda"),
             "target_size_as_tuple": torch.tensor([[1024., 1024.]] * 2, device="cuda")}
    with torch.no_grad():
        ref = ocond.conditioner(sd, batch)
    e1, e2 = rel_l2(c["crossattn"], ref["crossattn"]), rel_l2(c["vector"], ref["vector"])
    print(f"conditioner full depth crossattn rel-L2 {e1:.3e}, vector rel-L2 {e2:.3e}")
    assert c["crossattn"].shape == (2, 77, 2048) and c["vector"].shape == (2, 2816) and c["control"] is z
    assert e1 < 2e-2 and e2 < 2e-2
    assert uc["crossattn"].shape == (2, 77, 2048)
