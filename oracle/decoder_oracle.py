"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the temporal attention of the blurry-video decoder.

PARITY UNPINNED.  The block under test is /root/reference/model_variants/video_decoder.py:237-248 (AttnUpDecoderBlock2D.forward) and
:394-406 (UNetMidBlock2D.forward): reshape / rearrange to `(b h w) t c`, `temp_attn(...)`, rearrange back, blend with the scalar
`weight`.  `temp_attn` is `diffusers.models.attention_processor.Attention` (video_decoder.py:2; constructed at :204-216 and :353-365 with
norm_num_groups = resnet_groups, residual_connection = True, bias = True, upcast_softmax = True, rescale_output_factor =
output_scale_factor, _from_deprecated_attn_block = True).  diffusers (training env pin `diffusers==0.23.0`, requirements.txt:1) is not
installed in this image and not vendored in the reference tree, and the reference holds no fixture for this block -- so the Attention
arithmetic below is a RESTATEMENT of the published diffusers 0.23 algorithm (Attention.__init__ + AttnProcessor2_0.__call__ for a 3-D
input, no encoder states, no mask, no LoRA scale), anchored on the reference's own call site for everything around it:

    residual = hs                                                         # [B', t, c]
    hs = group_norm(hs.transpose(1, 2)).transpose(1, 2)                   # GroupNorm(32, c, eps, affine) over [B', c, t]
    q, k, v = to_q(hs), to_k(hs), to_v(hs)                                # Linear(c, c, bias=True)
    q, k, v -> [B', heads, t, d_h];  o = F.scaled_dot_product_attention(q, k, v)     # scale d_h^-1/2, softmax over t
    o -> [B', t, c];  o = to_out[0](o)                                    # Linear(c, c, bias=True); to_out[1] = Dropout(0)
    o = (o + residual) / rescale_output_factor                            # residual_connection = True
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as TF


@dataclass(frozen=True)
class DecoderAttnConfig:
    channels: int
    heads: int = 1                       # out_channels // attention_head_dim with attention_head_dim = out_channels (video_decoder.py:468,492)
    groups: int = 32                     # resnet_groups
    eps: float = 1e-6                    # resnet_eps (video_decoder.py:463,486)
    rescale_output_factor: float = 1.0   # output_scale_factor


def param_shapes(cfg: DecoderAttnConfig):
    C = cfg.channels
    return {"group_norm.weight": (C,), "group_norm.bias": (C,), "to_q.weight": (C, C), "to_q.bias": (C,), "to_k.weight": (C, C), "to_k.bias": (C,),
            "to_v.weight": (C, C), "to_v.bias": (C,), "to_out.0.weight": (C, C), "to_out.0.bias": (C,)}


def make_params(cfg: DecoderAttnConfig, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    out = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + 104729 + idx)
        if name == "group_norm.weight":
            p = 1.0 + 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        elif name == "group_norm.bias":
            p = 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        else:
            p = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(cfg.channels)
        out[name] = p.to(dtype)
    return out


def attention_forward(p: Dict[str, torch.Tensor], hs: torch.Tensor, cfg: DecoderAttnConfig) -> torch.Tensor:
    """diffusers Attention (AttnProcessor2_0) on [B', t, c] -- restated, see the module docstring."""
    residual = hs
    Bp, t, C = hs.shape
    hs = TF.group_norm(hs.transpose(1, 2), cfg.groups, p["group_norm.weight"], p["group_norm.bias"], cfg.eps).transpose(1, 2)
    q = TF.linear(hs, p["to_q.weight"], p["to_q.bias"])
    k = TF.linear(hs, p["to_k.weight"], p["to_k.bias"])
    v = TF.linear(hs, p["to_v.weight"], p["to_v.bias"])
    dh = C // cfg.heads
    q, k, v = (z.view(Bp, t, cfg.heads, dh).transpose(1, 2) for z in (q, k, v))
    o = TF.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(Bp, t, C)
    o = TF.linear(o, p["to_out.0.weight"], p["to_out.0.bias"])
    return (o + residual) / cfg.rescale_output_factor


def temporal_blend_reference_order(p: Dict[str, torch.Tensor], x: torch.Tensor, weight: float, time: int, cfg: DecoderAttnConfig) -> torch.Tensor:
    """video_decoder.py:241-248 / :398-406, same op order.  x: [(b t), c, h, w] -> same shape."""
    bt, c, h, w = x.shape
    b = bt // time
    res = x.reshape(b, time, c, h, w)
    res = res.permute(0, 3, 4, 1, 2).reshape(b * h * w, time, c)                         # 'b t c h w -> (b h w) t c'
    res = attention_forward(p, res, cfg).reshape(b, h, w, time, c)
    res = res.permute(0, 3, 4, 1, 2).reshape(b * time, c, h, w)                          # 'b h w t c -> (b t) c h w'
    return weight * x + (1 - weight) * res
