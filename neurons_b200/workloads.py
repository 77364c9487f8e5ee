"""Shapes of the motion-module calls the reference makes, and their algorithmic work.

One UNet3DConditionModel.forward with the inference-v3 config makes exactly 20 motion-module calls
(/root/reference/animatediff/models/unet.py:157,183,236 + unet_blocks.py:275,411,511,661,754; verified by instantiating
the reference), in this order of (channels, latent side) for a side-L latent:
    (320,L)x2 (640,L/2)x2 (1280,L/4)x2 (1280,L/8)x2 | (1280,L/8)x3 (1280,L/4)x3 (640,L/2)x3 (320,L)x3
The SparseCtrl ControlNet adds 8 calls with one attention block and max_len 32
(configs/inference/sparsectrl/latent_condition.yaml:11-17).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List


@dataclass(frozen=True)
class Call:
    channels: int
    side: int
    attn_blocks: int = 2
    max_len: int = 24


def unet_step_calls(latent_side: int) -> List[Call]:
    L = latent_side
    down = [(320, L)] * 2 + [(640, L // 2)] * 2 + [(1280, L // 4)] * 2 + [(1280, L // 8)] * 2
    up = [(1280, L // 8)] * 3 + [(1280, L // 4)] * 3 + [(640, L // 2)] * 3 + [(320, L)] * 3
    return [Call(c, s) for c, s in down + up]


def controlnet_step_calls(latent_side: int) -> List[Call]:
    L = latent_side
    return [Call(c, s, 1, 32) for c, s in [(320, L)] * 2 + [(640, L // 2)] * 2 + [(1280, L // 4)] * 2 + [(1280, L // 8)] * 2]


def module_flops(channels: int, tokens: int, frames: int, attn_blocks: int = 2, layers: int = 1) -> float:
    """Algorithmic FLOPs of one forward: 2*N*C^2*(2 + L*(4A+12)) + L*A*4*N*F*C."""
    C, N = channels, tokens
    return 2.0 * N * C * C * (2 + layers * (4 * attn_blocks + 12)) + layers * attn_blocks * 4.0 * N * frames * C


def module_min_bytes(channels: int, tokens: int, elem_size: int, attn_blocks: int = 2, layers: int = 1) -> float:
    """Whole-module HBM floor: read x twice (two-pass GroupNorm), write y once, read the parameters once."""
    C = channels
    params = C * C * (2 + layers * (4 * attn_blocks + 12)) + C * (6 + layers * (3 * attn_blocks + 12))
    return 3.0 * tokens * C * elem_size + params * elem_size


def step_flops(calls: List[Call], batch: int, frames: int) -> float:
    return sum(module_flops(c.channels, batch * frames * c.side * c.side, frames, c.attn_blocks) for c in calls)
