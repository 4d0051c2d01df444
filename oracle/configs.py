"""ORACLE (test infrastructure) — architecture hyper-parameters of the reference configs.

STAGE2_*  : model_configs/juggernautXL.yaml:21-67 (control_stage_config / network_config params)
SR3_UNET  : configs/sr_sr3.json:41-57 + models/sr3_model/networks.py:84-136 (image_size 224)
The *_TEST variants keep every channel width (LightGLVUNet hard-codes the adapter widths for
"XL-base", SR_modules.py:544-549) and only reduce transformer depth so CPU tests build quickly.
"""
import copy

_COMMON = dict(
    adm_in_channels=2816, num_classes="sequential", use_checkpoint=False, in_channels=4, out_channels=4,
    model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4],
    num_head_channels=64, use_spatial_transformer=True, use_linear_in_transformer=True,
    transformer_depth=[1, 2, 10], context_dim=2048, spatial_transformer_attn_type="softmax-xformers", legacy=False,
)
STAGE2_CONTROL = dict(_COMMON, input_upscale=1)
STAGE2_UNET = dict(_COMMON, mode="XL-base", project_type="ZeroSFT", project_channel_scale=2)


def reduced(cfg: dict, depth=(1, 1, 1)) -> dict:
    c = copy.deepcopy(cfg)
    c["transformer_depth"] = list(depth)
    return c


STAGE2_CONTROL_TEST = reduced(STAGE2_CONTROL)
STAGE2_UNET_TEST = reduced(STAGE2_UNET)

SR3_UNET = dict(in_channel=6, out_channel=3, norm_groups=32, inner_channel=64, channel_mults=[1, 2, 4, 8, 8],
                attn_res=[28], res_blocks=1, dropout=0.2, image_size=224)
SR3_SCHEDULE = dict(n_timestep=50, linear_start=1e-6, linear_end=1e-2)  # BASELINE config 1 (reference default: 500)

# first stage (SDXL VAE): model_configs/juggernautXL.yaml:107-125.  attn_type "vanilla-xformers" and "vanilla" compute the
# same single-head attention (model.py:158-263); the oracle / golden side uses "vanilla" (no xformers in this image).
VAE_DDCONFIG = dict(attn_type="vanilla", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
VAE_EMBED_DIM = 4
