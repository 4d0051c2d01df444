"""ORACLE (test infrastructure) — restatement of the output side: wavelet colour fix (utils/colorfix.py:73-119) and
Tensor2PIL (models/util.py:159-166), plain torch.  Pinned against the real reference functions in
oracle/make_golden.py (colorfix()); golden outputs in tests/golden/colorfix_48.pt."""
import torch
import torch.nn.functional as F


def wavelet_blur(image, radius):
    """colorfix.py:73-91."""
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]],
                     dtype=image.dtype, device=image.device)[None, None].repeat(3, 1, 1, 1)
    image = F.pad(image, (radius, radius, radius, radius), mode="replicate")
    return F.conv2d(image, k, groups=3, dilation=radius)


def wavelet_decomposition(image, levels=5):
    """colorfix.py:93-106."""
    high = torch.zeros_like(image)
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high += image - low
        image = low
    return high, image


def wavelet_reconstruction(content, style):
    """colorfix.py:108-119."""
    return wavelet_decomposition(content)[0] + wavelet_decomposition(style)[1]


def tensor_to_uint8(x, h0, w0):
    """Tensor2PIL's array — models/util.py:159-166."""
    y = F.interpolate(x.unsqueeze(0), size=(h0, w0), mode="bicubic")
    return torch.from_numpy((y.squeeze(0).permute(1, 2, 0) * 127.5 + 127.5).cpu().numpy().clip(0, 255).astype("uint8"))
