"""ORACLE (test infrastructure) — deterministic random weights keyed by parameter *name*.

The reference zero-initialises 73 tensors (every ResBlock conv2, SpatialTransformer.proj_out, the
UNet `out` conv, ZeroSFT zero_mul / zero_add / zero_conv, the control hint conv: util.py:233-239).
With those left at zero the random-init network outputs exactly 0 and parity would be vacuous, so
all tensors are re-randomised.  Each tensor is drawn from its own generator seeded by
(seed, crc32(name)), so the reference modules (build container), the oracle and the product
modules (GPU box) get bit-identical weights independent of construction order.
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch


# Output layers of residual branches.  The reference zero-initialises these in stage 2 (ResBlock conv2,
# SpatialTransformer.proj_out, ZeroSFT zero_mul / zero_add / zero_conv, the control hint conv, UNet.out);
# they are re-randomised at 1/4 of the default scale so the random network stays near the identity-
# residual regime it is initialised in (at full scale a random-weight SDXL UNet amplifies the
# unavoidable bf16 operand rounding of every layer to ~1e-2 on its own, leaving no room to judge the
# kernels).  The SR3 residual ends (block2 conv, attention out conv) get the same treatment.
BRANCH_END = (".out_layers.3.", ".proj_out.", ".zero_mul.", ".zero_add.", ".zero_conv.", ".input_hint_block.0.",
              ".out.2.", ".block2.block.3.", ".attn.out.", ".conv2.",   # .conv2. = the first stage's ResnetBlock branch end
              ".out_proj.", ".fc2.", ".c_proj.")                        # text towers: attention / MLP output projections
BRANCH_END_GAIN = 0.25


def fill_(sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """In-place: weights ~ N(0, 1/(3 fan_in)) (the std of torch's default kaiming-uniform init; x 1/4 for
    BRANCH_END tensors), biases ~ N(0, 0.02^2), norm scales ~ 1 + N(0, 0.05^2), norm shifts ~ N(0, 0.05^2)."""
    for name in sorted(sd.keys()):
        t = sd[name]
        if not t.is_floating_point():
            continue
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**63))
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        dotted = "." + name
        is_norm = t.dim() == 1 and any(
            s in dotted for s in ("norm", ".in_layers.0.", ".out_layers.0.", ".out.0.", ".block.0.", ".ln_"))   # incl. norm1/norm2/norm_out, ln_1/ln_2/ln_final
        if is_norm:
            r = r * 0.05 + (1.0 if leaf == "weight" else 0.0)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            r = r * (1.0 / (3.0 * fan_in)) ** 0.5
            if any(b in dotted for b in BRANCH_END):
                r = r * BRANCH_END_GAIN
        else:
            r = r * 0.02
        with torch.no_grad():
            t.copy_(r.to(t.dtype))
    return sd
