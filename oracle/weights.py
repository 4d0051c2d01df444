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


def fill_(sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """In-place: weights ~ N(0, 1/(3 fan_in)) (the std of torch's default kaiming-uniform init),
    biases ~ N(0, 0.02^2), norm scales ~ 1 + N(0, 0.05^2), norm shifts ~ N(0, 0.05^2)."""
    for name in sorted(sd.keys()):
        t = sd[name]
        if not t.is_floating_point():
            continue
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**63))
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        dotted = "." + name
        is_norm = t.dim() == 1 and any(
            s in dotted for s in ("norm", ".in_layers.0.", ".out_layers.0.", ".out.0.", ".block.0."))
        if is_norm:
            r = r * 0.05 + (1.0 if leaf == "weight" else 0.0)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            r = r * (1.0 / (3.0 * fan_in)) ** 0.5
        else:
            r = r * 0.02
        with torch.no_grad():
            t.copy_(r.to(t.dtype))
    return sd
