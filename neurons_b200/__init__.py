"""neurons_b200 -- B200-native (sm_100a) drop-in for the AnimateDiff motion module of xmed-lab/NEURONS.

Public surface (mirrors /root/reference/animatediff/models/motion_module.py):
    get_motion_module, VanillaTemporalModule      construct modules with the reference's signature / checkpoint layout
    patch(model), invalidate(model)                rebind forward on reference modules already inside a UNet / ControlNet
    torch.ops.neurons_mm.forward                   the custom op (CUDA dispatch key only)
    InflatedGroupNorm, patch_group_norms           the per-frame GroupNorm of ResnetBlock3D either side of the module (resnet.py:21-29)
    patch(model, carry_stats=True)                 ... whose statistics then come from the motion module's last kernel (attach_sums / carried_sums)
    Transformer3DModel, patch_spatial(model)      the spatial transformer that precedes the motion module in every CrossAttn block
                                                   (attention.py:31-300): same kernels + a flash-style spatial attention
    temporal_attention_blend, patch_video_decoder, decode_latents_batched   the blurry-video decoder's t = 6 temporal attention + blend and the
                                                   batched VAE decode (video_decoder.py:237-248,394-406; SURVEY 8(f) N4, parity unpinned)
The arithmetic lives in libneurons_mm.so (C ABI: include/neurons_mm.h), built by `python -m neurons_b200.build`.
"""
from .lib import NmmError, launch_count, load as load_library          # noqa: F401
from .ops import ModuleConfig                                           # noqa: F401
from .motion_module import (VanillaTemporalModule, attach_sums, carried_sums, config_of, get_motion_module, invalidate,   # noqa: F401
                            motion_forward, patch, zero_module)
from .resnet_norm import InflatedGroupNorm, patch_group_norms            # noqa: F401
from .video_decoder import decode_latents_batched, patch_video_decoder, temporal_attention_blend          # noqa: F401
from .spatial_transformer import (SpatialConfig, Transformer3DModel, patch_spatial, spatial_attention, spatial_config_of,   # noqa: F401
                                  spatial_forward)

__all__ = ["VanillaTemporalModule", "get_motion_module", "patch", "invalidate", "motion_forward", "zero_module", "config_of",
           "ModuleConfig", "NmmError", "launch_count", "load_library", "InflatedGroupNorm", "patch_group_norms", "attach_sums", "carried_sums",
           "decode_latents_batched", "patch_video_decoder", "temporal_attention_blend", "SpatialConfig", "Transformer3DModel", "patch_spatial", "spatial_attention", "spatial_config_of", "spatial_forward"]
