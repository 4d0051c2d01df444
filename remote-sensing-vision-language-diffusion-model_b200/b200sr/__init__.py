"""b200sr — B200-native (sm_100a) diffusion-denoiser hot path of
Bluear7878/Remote-Sensing-Vision-Language-Diffusion-Model behind the reference's nn.Module API."""
__version__ = "0.1.0"
