"""ORACLE (test infrastructure) — seeded synthetic inputs shared by golden generation, parity
tests, smoke() and bench.py (BASELINE.json: LLaVA captions / CLIP embeddings are replaced by a
fixed synthetic text embedding)."""
import torch


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def stage2_inputs(latent: int = 128, seed: int = 1234, batch: int = 1):
    """x (already scaled noise), control (LQ latent), c / uc dicts: crossattn [B,77,2048], vector [B,2816]."""
    g = _gen(seed)
    x = torch.randn(batch, 4, latent, latent, generator=g) * (1.0 + 14.6146**2) ** 0.5
    control = torch.randn(batch, 4, latent, latent, generator=g)
    c = {"crossattn": torch.randn(batch, 77, 2048, generator=g), "vector": torch.randn(batch, 2816, generator=g),
         "control": control}
    uc = {"crossattn": torch.randn(batch, 77, 2048, generator=g), "vector": torch.randn(batch, 2816, generator=g),
          "control": control}
    return x, c, uc


def sr3_inputs(size: int = 128, seed: int = 0, steps: int = 50):
    """cond (bicubic-upsampled LR in [-1,1]) and the per-step noises (noises[0] = initial image)."""
    g = _gen(seed)
    lr = torch.rand(1, 3, size // 8, size // 8, generator=g) * 2 - 1
    cond = torch.nn.functional.interpolate(lr, scale_factor=8, mode="bicubic", align_corners=False).clamp(-1, 1)
    noises = [torch.randn(1, 3, size, size, generator=g) for _ in range(steps + 1)]
    return cond, noises
