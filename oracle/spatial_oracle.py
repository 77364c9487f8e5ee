"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the UNet's spatial Transformer3DModel forward.

The checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import it.

Restates, in plain torch-on-CPU, /root/reference/animatediff/models/attention.py:
    Transformer3DModel.forward            :95-148
    BasicTransformerBlock.forward         :258-300   (NEURONS configuration: LayerNorm norms, attn1 = plain CrossAttention self-attention
                                                      -- unet_use_cross_frame_attention False --, attn2 = text cross-attention, GEGLU
                                                      feed-forward, no temporal attention inside the block, no attention mask)
with the inherited diffusers-0.11.1 CrossAttention / FeedForward / GEGLU arithmetic read from the in-tree copy
animatediff/models/motion_module_new.py (CrossAttention.forward :194-256, _attention :258-287, head reshapes :181-193,
FeedForward :429-471, GEGLU :497-518), because diffusers is not vendored.

PARITY PIN: no golden vectors exist in the reference for this path either; the pin is the unmodified reference
class imported through oracle/unet_shim.py -- tests/test_spatial_oracle.py::test_pin_against_live_reference -- and the fixtures
tests/golden/sp_*.pt that oracle/gen_spatial_golden.py produced from it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as TF

GN_GROUPS = 32      # attention.py:37 norm_num_groups default, unet_blocks passes resnet_groups = 32
GN_EPS = 1e-6       # attention.py:60
LN_EPS = 1e-5       # nn.LayerNorm default (attention.py:196,214,221)


@dataclass(frozen=True)
class SpatialConfig:
    channels: int
    heads: int = 8              # attention_head_dim=8 in the UNet config means 8 HEADS (unet_blocks.py:229: attn_num_head_channels)
    layers: int = 1             # num_layers
    ctx_dim: int = 768          # cross_attention_dim
    conv_proj: bool = True      # use_linear_projection = False: proj_in / proj_out are 1x1 convolutions

    @property
    def head_dim(self) -> int:
        return self.channels // self.heads


def param_shapes(cfg: SpatialConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of the reference Transformer3DModel (listed from the live class)."""
    C, D = cfg.channels, cfg.ctx_dim
    pw = (C, C, 1, 1) if cfg.conv_proj else (C, C)
    s: Dict[str, Tuple[int, ...]] = {"norm.weight": (C,), "norm.bias": (C,), "proj_in.weight": pw, "proj_in.bias": (C,)}
    for l in range(cfg.layers):
        b = f"transformer_blocks.{l}."
        s[b + "attn1.to_q.weight"] = (C, C); s[b + "attn1.to_k.weight"] = (C, C); s[b + "attn1.to_v.weight"] = (C, C)
        s[b + "attn1.to_out.0.weight"] = (C, C); s[b + "attn1.to_out.0.bias"] = (C,)
        s[b + "norm1.weight"] = (C,); s[b + "norm1.bias"] = (C,)
        s[b + "attn2.to_q.weight"] = (C, C); s[b + "attn2.to_k.weight"] = (C, D); s[b + "attn2.to_v.weight"] = (C, D)
        s[b + "attn2.to_out.0.weight"] = (C, C); s[b + "attn2.to_out.0.bias"] = (C,)
        s[b + "norm2.weight"] = (C,); s[b + "norm2.bias"] = (C,)
        s[b + "ff.net.0.proj.weight"] = (8 * C, C); s[b + "ff.net.0.proj.bias"] = (8 * C,)
        s[b + "ff.net.2.weight"] = (C, 4 * C); s[b + "ff.net.2.bias"] = (C,)
        s[b + "norm3.weight"] = (C,); s[b + "norm3.bias"] = (C,)
    s["proj_out.weight"] = pw
    s["proj_out.bias"] = (C,)
    return s


def make_params(cfg: SpatialConfig, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic name-keyed weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear / conv, norm affines perturbed from (1, 0)."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + 7919 + idx)
        is_norm = "norm" in name.split(".")[-2]
        if is_norm and name.endswith("weight"):
            p = 1.0 + 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        elif is_norm:
            p = 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        else:
            fan_in = shape[1] if len(shape) >= 2 else (4 * cfg.channels if name.endswith("net.2.bias") else cfg.channels)
            p = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(fan_in)
        out[name] = p.to(dtype)
    return out


def make_inputs(cfg: SpatialConfig, batch: int, frames: int, height: int, width: int, ctx_len: int, seed: int, dtype=torch.float32,
                layout: str = "bcfhw"):
    """x: seeded N(0,1) [B, C, F, H, W] (contiguous = what ResnetBlock3D hands over; "bfchw" = [B,F,C,H,W]-storage view);
    encoder_hidden_states: N(0,1) [B, ctx_len, ctx_dim] (CLIP text states are O(1) after its final LayerNorm)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, cfg.channels, frames, height, width), generator=g, dtype=torch.float64).to(dtype)
    if layout == "bfchw":
        x = x.permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4)
    ctx = torch.randn((batch, ctx_len, cfg.ctx_dim), generator=g, dtype=torch.float64).to(dtype)
    return x, ctx


def _heads_to_batch(t: torch.Tensor, heads: int) -> torch.Tensor:
    """motion_module_new.py:181-186  [b, s, h*d] -> [b*h, s, d]"""
    b, s, d = t.shape
    return t.reshape(b, s, heads, d // heads).permute(0, 2, 1, 3).reshape(b * heads, s, d // heads)


def _batch_to_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    """motion_module_new.py:188-193  [b*h, s, d] -> [b, s, h*d]"""
    bh, s, d = t.shape
    return t.reshape(bh // heads, heads, s, d).permute(0, 2, 1, 3).reshape(bh // heads, s, d * heads)


def _attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
    """CrossAttention._attention, motion_module_new.py:258-287 (no mask; softmax in the input dtype)."""
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device), q, k.transpose(-1, -2), beta=0, alpha=scale)
    probs = scores.softmax(dim=-1)
    return torch.bmm(probs, v)


def _cross_attention(p: Dict[str, torch.Tensor], prefix: str, hidden: torch.Tensor, context, heads: int) -> torch.Tensor:
    """CrossAttention.forward, motion_module_new.py:194-256 (group_norm None, added_kv_proj_dim None, no slicing)."""
    q = TF.linear(hidden, p[prefix + "to_q.weight"])
    ctx = hidden if context is None else context
    k = TF.linear(ctx, p[prefix + "to_k.weight"])
    v = TF.linear(ctx, p[prefix + "to_v.weight"])
    dh = q.shape[-1] // heads
    o = _attention(_heads_to_batch(q, heads), _heads_to_batch(k, heads), _heads_to_batch(v, heads), dh ** -0.5)
    o = _batch_to_heads(o, heads)
    return TF.linear(o, p[prefix + "to_out.0.weight"], p[prefix + "to_out.0.bias"])      # to_out[1] is Dropout(0)


def _feed_forward(p: Dict[str, torch.Tensor], prefix: str, hidden: torch.Tensor) -> torch.Tensor:
    """FeedForward + GEGLU, motion_module_new.py:441-471,497-518: proj -> chunk(2) -> value * gelu(gate) -> Linear."""
    u = TF.linear(hidden, p[prefix + "net.0.proj.weight"], p[prefix + "net.0.proj.bias"])
    a, gate = u.chunk(2, dim=-1)
    return TF.linear(a * TF.gelu(gate), p[prefix + "net.2.weight"], p[prefix + "net.2.bias"])


def forward_reference_order(params: Dict[str, torch.Tensor], x: torch.Tensor, encoder_hidden_states: torch.Tensor, cfg: SpatialConfig) -> torch.Tensor:
    """Transformer3DModel.forward, attention.py:95-148, same op order.  Returns logical [B, C, F, H, W] (a view over [B,F,C,H,W] storage)."""
    assert x.dim() == 5                                                              # :97
    B, C, F, H, W = x.shape
    hs = x.permute(0, 2, 1, 3, 4).reshape(B * F, C, H, W)                            # :99  "b c f h w -> (b f) c h w"
    ehs = encoder_hidden_states.repeat_interleave(F, dim=0)                          # :100 "b n c -> (b f) n c"
    residual = hs
    hs = TF.group_norm(hs, GN_GROUPS, params["norm.weight"], params["norm.bias"], GN_EPS)      # :106
    if cfg.conv_proj:                                                                # :107-110
        hs = TF.conv2d(hs, params["proj_in.weight"], params["proj_in.bias"])
        hs = hs.permute(0, 2, 3, 1).reshape(B * F, H * W, C)
    else:                                                                            # :111-114
        hs = hs.permute(0, 2, 3, 1).reshape(B * F, H * W, C)
        hs = TF.linear(hs, params["proj_in.weight"], params["proj_in.bias"])
    for l in range(cfg.layers):                                                      # :117-123 -> BasicTransformerBlock.forward :258-300
        b = f"transformer_blocks.{l}."
        n = TF.layer_norm(hs, (C,), params[b + "norm1.weight"], params[b + "norm1.bias"], LN_EPS)          # :260-262
        hs = _cross_attention(params, b + "attn1.", n, None, cfg.heads) + hs                                 # :277-280
        n = TF.layer_norm(hs, (C,), params[b + "norm2.weight"], params[b + "norm2.bias"], LN_EPS)          # :284-286
        hs = _cross_attention(params, b + "attn2.", n, ehs, cfg.heads) + hs                                  # :287-292
        n = TF.layer_norm(hs, (C,), params[b + "norm3.weight"], params[b + "norm3.bias"], LN_EPS)
        hs = _feed_forward(params, b + "ff.", n) + hs                                                        # :295
    if cfg.conv_proj:                                                                # :126-130
        hs = hs.reshape(B * F, H, W, C).permute(0, 3, 1, 2).contiguous()
        hs = TF.conv2d(hs, params["proj_out.weight"], params["proj_out.bias"])
    else:                                                                            # :131-135
        hs = TF.linear(hs, params["proj_out.weight"], params["proj_out.bias"])
        hs = hs.reshape(B * F, H, W, C).permute(0, 3, 1, 2).contiguous()
    out = hs + residual                                                              # :137
    return out.reshape(B, F, C, H, W).permute(0, 2, 1, 3, 4)                         # :139 "(b f) c h w -> b c f h w"


def flops(cfg: SpatialConfig, batch: int, frames: int, positions: int, ctx_len: int) -> float:
    """Algorithmic FLOPs of one call: Linear layers 2*M*N*K + both attentions 4*N*L*C."""
    C, D = cfg.channels, cfg.ctx_dim
    N = batch * frames * positions
    per_layer = 2.0 * N * C * C * (3 + 1 + 1 + 1 + 8 + 4) + 2.0 * batch * ctx_len * 2 * C * D + 4.0 * N * positions * C + 4.0 * N * ctx_len * C
    return 2.0 * 2.0 * N * C * C + cfg.layers * per_layer
