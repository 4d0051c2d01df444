"""Output side of the pipeline on the sm_100a kernels: wavelet colour fix and the
tensor -> uint8 image conversion.

Mirrored reference code (relative to the reference root), same function names and argument meaning:
  wavelet_blur / wavelet_decomposition / wavelet_reconstruction      utils/colorfix.py:73-119
  Tensor2PIL                                                          models/util.py:159-166
(color_fix_type defaults to "Wavelet" in infer.py / infer_dir.py; the AdaIN variant, colorfix.py:44-71, is not selected
by the shipped drivers and is not built.)
"""
from __future__ import annotations

import torch

from . import ops


def wavelet_blur(image: torch.Tensor, radius: int) -> torch.Tensor:
    """colorfix.py:73-91: depthwise 3x3 binomial blur with dilation `radius`, replicate padding."""
    return ops.wavelet_level(image.float().contiguous(), radius)


def wavelet_decomposition(image: torch.Tensor, levels: int = 5):
    """colorfix.py:93-106: returns (high_freq, low_freq)."""
    image = image.float().contiguous()
    high = torch.empty_like(image)
    for i in range(levels):
        image = ops.wavelet_level(image, 2 ** i, high=high, first=(i == 0))
    return high, image


def wavelet_reconstruction(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    """colorfix.py:108-119: the content's high frequencies on the style's low frequencies."""
    content_high, _ = wavelet_decomposition(content_feat)
    style = style_feat.float().contiguous()
    for i in range(5):
        style = ops.wavelet_level(style, 2 ** i)
    return ops.add_f32(content_high, style)


def tensor_to_uint8(x: torch.Tensor, h0: int, w0: int) -> torch.Tensor:
    """Tensor2PIL without the PIL object (models/util.py:159-166): [C, H, W] in [-1, 1] -> uint8 [h0, w0, C] on the
    device (bicubic resize, * 127.5 + 127.5, clip, truncate); ``PIL.Image.fromarray(result.cpu().numpy())`` is the image."""
    return ops.image_to_u8(x.float().contiguous(), h0, w0)
