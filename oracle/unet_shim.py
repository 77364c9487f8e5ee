"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference UNet3DConditionModel / SparseControlNetModel.

SURVEY.md 8(c) "shim for the whole UNet": the reference files animatediff/models/{unet,unet_blocks,attention,resnet,
sparse_controlnet}.py import a handful of `diffusers` (0.11.1) symbols that are not installed in this image.  This module registers
stand-ins for exactly those symbols in sys.modules and then imports the reference package from /root/reference (a namespace package:
no __init__.py).  No reference source is copied or modified; the stand-ins restate only glue that sits OUTSIDE the compared region:

    diffusers.configuration_utils.{ConfigMixin, register_to_config}   -> bound ctor args kept in `.config`
    diffusers.modeling_utils.ModelMixin                              -> nn.Module with .dtype / .device
    diffusers.utils.{BaseOutput, logging}, diffusers.utils.import_utils.is_xformers_available
    diffusers.models.embeddings.{Timesteps, TimestepEmbedding}       -> sinusoidal timestep features (flip_sin_to_cos, freq_shift) and
                                                                        linear_1 -> SiLU -> linear_2  (shared by both sides of any A/B)
    diffusers.models.attention.{CrossAttention, FeedForward}         -> the reference's own in-tree copy (motion_module_new.py)
    diffusers.models.attention.AdaLayerNorm, diffusers.models.unet_2d_condition.UNet2DConditionModel -> placeholders (never instantiated)

Used by tests/test_reference_unet.py (patch() on reference-built models) and oracle/gen_unet_golden.py (activation fixtures).
It never travels to the GPU box (the reference tree is absent there) and nothing in neurons_b200/ may import it.
"""
from __future__ import annotations

import functools
import importlib
import inspect
import math
import sys
import types

import torch
from torch import nn

from . import ref_shim

_loaded = None


class _Config(dict):
    __getattr__ = dict.__getitem__


def _register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        init(self, *args, **kwargs)
        object.__setattr__(self, "_shim_config", _Config(cfg))
    return wrapper


class _ConfigMixin:
    @property
    def config(self):
        return self._shim_config


class _ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _Timesteps(nn.Module):
    """Sinusoidal timestep features (diffusers get_timestep_embedding: max_period 1e4, scale 1)."""

    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class _TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim: int = None, **_):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


class _Logger:
    def __getattr__(self, name):
        return lambda *a, **k: None


def available() -> bool:
    return ref_shim.available()


def load():
    """-> namespace with UNet3DConditionModel, SparseControlNetModel and the reference's motion_module python module."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = ref_shim.reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set NEURONS_REF or mount /root/reference)")
    mm = ref_shim.load_reference_motion_module()          # registers the base diffusers stubs + CrossAttention / FeedForward
    d, du = sys.modules["diffusers"], sys.modules["diffusers.utils"]
    dma = sys.modules["diffusers.models.attention"]

    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m
    cu = mod("diffusers.configuration_utils"); cu.ConfigMixin = _ConfigMixin; cu.register_to_config = _register_to_config
    mu = mod("diffusers.modeling_utils"); mu.ModelMixin = _ModelMixin
    lg = mod("diffusers.utils.logging"); lg.get_logger = lambda *a, **k: _Logger()
    du.logging = lg
    emb = mod("diffusers.models.embeddings"); emb.Timesteps = _Timesteps; emb.TimestepEmbedding = _TimestepEmbedding
    u2d = mod("diffusers.models.unet_2d_condition"); u2d.UNet2DConditionModel = object
    dma.AdaLayerNorm = type("AdaLayerNorm", (nn.Module,), {})
    d.configuration_utils, d.modeling_utils = cu, mu
    sys.modules["diffusers.models"].embeddings = emb
    sys.modules["diffusers.models"].unet_2d_condition = u2d
    if root not in sys.path:
        sys.path.insert(0, root)
    # the package's motion_module must be the SAME module object ref_shim loaded (one class identity for isinstance / patch)
    sys.modules.setdefault("animatediff.models.motion_module", mm)
    unet = importlib.import_module("animatediff.models.unet")
    cn = importlib.import_module("animatediff.models.sparse_controlnet")
    _loaded = types.SimpleNamespace(UNet3DConditionModel=unet.UNet3DConditionModel, SparseControlNetModel=cn.SparseControlNetModel,
                                    motion_module=sys.modules["animatediff.models.motion_module"], unet_module=unet, controlnet_module=cn)
    return _loaded


# SD1.5 topology + configs/inference/inference-v3.yaml (UNet) and configs/inference/sparsectrl/latent_condition.yaml (ControlNet)
UNET_KW = dict(sample_size=64, in_channels=4, out_channels=4, cross_attention_dim=768, attention_head_dim=8,
               block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
               down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
               up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
               use_inflated_groupnorm=True, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
               motion_module_type="Vanilla",
               motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                                         temporal_position_encoding=True, temporal_attention_dim_div=1, zero_initialize=True))
CONTROLNET_KW = dict(in_channels=4, cross_attention_dim=768, attention_head_dim=8, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                     down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                     set_noisy_sample_input_to_zero=True, use_simplified_condition_embedding=True, conditioning_channels=4,
                     use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False, motion_module_type="Vanilla",
                     motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self",),
                                               temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1))


def build_unet(seed: int = 0, **overrides):
    """Random-init reference UNet (SD1.5 topology, v3 motion modules); every motion proj_out re-randomised N(0, 0.02) (else identity)."""
    ns = load()
    torch.manual_seed(seed)
    kw = dict(UNET_KW)
    kw.update(overrides)
    unet = ns.UNet3DConditionModel(**kw).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in unet.modules():
            if type(m).__name__ == "VanillaTemporalModule":
                po = m.temporal_transformer.proj_out
                po.weight.copy_(torch.randn(po.weight.shape, generator=g) * 0.02)
                po.bias.copy_(torch.randn(po.bias.shape, generator=g) * 0.02)
    return unet


def build_controlnet(seed: int = 0, **overrides):
    ns = load()
    torch.manual_seed(seed)
    kw = dict(CONTROLNET_KW)
    kw.update(overrides)
    cn = ns.SparseControlNetModel(**kw).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in cn.modules():
            if type(m).__name__ == "VanillaTemporalModule":
                po = m.temporal_transformer.proj_out
                po.weight.copy_(torch.randn(po.weight.shape, generator=g) * 0.02)
                po.bias.copy_(torch.randn(po.bias.shape, generator=g) * 0.02)
    return cn
