"""InflatedGroupNorm of the reference's ResnetBlock3D, on the CUDA kernels.

Mirrors /root/reference/animatediff/models/resnet.py:21-29: an nn.GroupNorm applied per frame to a [b, c, f, h, w] tensor.
The reference rearranges to (b f) c h w (a copy), normalises, and rearranges back (another copy); the kernel reads x in its own
strides and writes the contiguous [b, c, f, h, w] result directly.  Same parameters / state_dict keys (`weight`, `bias`) as
nn.GroupNorm, so checkpoints load unchanged.  `patch_group_norms(model)` rebinds `forward` on the reference's own
InflatedGroupNorm instances (32 groups, affine) the same way `patch()` does for the motion modules.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from . import ops


def _gn_forward(self: nn.GroupNorm, x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 5:
        raise AssertionError(f"Expected a [b, c, f, h, w] tensor, got ndim={x.dim()}.")
    if not x.is_cuda:
        raise RuntimeError("neurons_b200.InflatedGroupNorm: CUDA tensors only (there is no CPU path)")
    if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
        raise RuntimeError("neurons_b200.InflatedGroupNorm is inference-only: call it under torch.no_grad()")
    # statistics emitted by the producer of x (a carry_stats-patched motion module): one pass over x instead of two
    from .motion_module import carried_sums
    return ops.inflated_groupnorm(x, self.weight, self.bias, self.eps, silu=False, sums=carried_sums(x))


class InflatedGroupNorm(nn.GroupNorm):
    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, affine: bool = True, **kw):
        if num_groups != 32 or not affine:
            raise NotImplementedError("neurons_b200.InflatedGroupNorm supports 32 affine groups (the NEURONS configuration)")
        super().__init__(num_groups, num_channels, eps=eps, affine=affine, **kw)

    forward = _gn_forward


def patch_group_norms(model: nn.Module) -> int:
    """Rebind forward on every InflatedGroupNorm (by class name, 32 groups, affine) inside `model`; returns how many."""
    n = 0
    for m in model.modules():
        if isinstance(m, nn.GroupNorm) and type(m).__name__ == "InflatedGroupNorm" and m.num_groups == 32 and m.affine:
            m.forward = types.MethodType(_gn_forward, m)
            n += 1
    return n
