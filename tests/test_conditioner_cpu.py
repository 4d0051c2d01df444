"""Conditioner on CPU: oracle vs the golden outputs of the installed transformers text
models, the drop-in modules' wiring through the test double, and parameter names of the third-party modules."""
import os

import torch

from oracle import conditioner as ocond, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _model(g):
    from b200sr import conditioner

    m = conditioner.GeneralConditionerWithControl(_clip_layers=g["clip_layers"], _clip_layer_idx=g["clip_idx"],
                                                  _bigg_layers=g["bigg_layers"]).eval()
    weights.fill_(m.state_dict(), 0)
    return m


def _batch(g):
    ids = g["ids"]
    return {"txt": (ids, ids.flip(0)), "original_size_as_tuple": torch.tensor([[1024., 1024.]] * 2),
            "crop_coords_top_left": torch.zeros(2, 2), "target_size_as_tuple": torch.tensor([[1024., 1024.]] * 2)}


def test_oracle_matches_transformers_golden():
    g = torch.load(os.path.join(GOLDEN, "conditioner_small.pt"), weights_only=False)
    sd = {k: v.detach() for k, v in _model(g).state_dict().items()}
    with torch.no_grad():
        hl = ocond.clip_l_hidden_states(sd, "embedders.0.transformer.text_model.", g["ids"])[g["clip_idx"]]
        pen, pooled = ocond.openclip_text(sd, "embedders.1.model.", g["ids"])
        out = ocond.conditioner(sd, _batch(g), g["clip_idx"])
    assert (hl - g["clip_hidden"]).abs().max().item() < 1e-5
    assert (pen - g["bigg_penultimate"]).abs().max().item() < 1e-5 and (pooled - g["bigg_pooled"]).abs().max().item() < 1e-5
    assert torch.allclose(out["crossattn"], g["crossattn"], atol=1e-5) and torch.allclose(out["vector"], g["vector"], atol=1e-5)
    assert out["crossattn"].shape == (2, 77, 2048) and out["vector"].shape == (2, 2816)


def test_parameter_names_follow_the_third_party_modules():
    import transformers

    from b200sr import conditioner

    m = conditioner.GeneralConditionerWithControl()
    cfg = transformers.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    with torch.device("meta"):
        hf = transformers.CLIPTextModel(cfg)
    ours = {k: tuple(v.shape) for k, v in m.embedders[0].transformer.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "position_ids" not in k}
    assert ours == theirs
    g = {k: tuple(v.shape) for k, v in m.embedders[1].model.state_dict().items()}
    assert g["transformer.resblocks.31.attn.in_proj_weight"] == (3840, 1280) and g["text_projection"] == (1280, 1280)
    assert g["positional_embedding"] == (77, 1280) and g["transformer.resblocks.0.mlp.c_fc.weight"] == (5120, 1280)


def test_module_wiring_against_golden(monkeypatch):
    import ops_double
    from b200sr import ops

    ops_double.install(monkeypatch, ops)
    g = torch.load(os.path.join(GOLDEN, "conditioner_small.pt"), weights_only=False)
    m = _model(g)
    out = m(_batch(g))
    assert out["crossattn"].shape == (2, 77, 2048) and out["vector"].shape == (2, 2816)
    assert rel_l2(out["crossattn"], g["crossattn"]) < 1e-2
    assert rel_l2(out["vector"], g["vector"]) < 1e-2
    c, uc = m.get_unconditional_conditioning(_batch(g), None, ["txt"])
    assert float(uc["crossattn"].abs().max()) == 0.0 and float(uc["vector"][:, :1280].abs().max()) == 0.0
    assert torch.equal(uc["vector"][:, 1280:], c["vector"][:, 1280:])
