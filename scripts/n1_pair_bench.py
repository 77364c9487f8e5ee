#!/usr/bin/env python
"""SURVEY 8(f) N1 with the real producer: a CrossAttn block's transformer pair (spatial Transformer3DModel -> motion module,
unet_blocks.py:409-411) per UNet level, bf16, config-2 shapes, in-stream time per pair:
    A  no statistics carried                      (spatial: gn_stats; motion: gn_stats)
    B' motion module emits the sums of its output (for the next ResnetBlock3D.norm1), computes those of its input itself
    B  + the spatial transformer's proj_out emits the sums of ITS output and the motion module takes them (no statistics pass over x)
B' - B is what this fusion saves per block; B' - A what the motion module's own emission costs."""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    dev = torch.device("cuda", 0)
    kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
    for C, side in ((320, 64), (640, 32), (1280, 16)):
        with torch.no_grad():
            holders, xs = [], []
            for _ in range(2):            # two buffer sets: inputs larger than what one call leaves in L2 at the big levels
                h = nn.Module()
                with torch.device(dev):
                    h.sp = nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=C // 8, in_channels=C, cross_attention_dim=768,
                                                 unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
                    h.mm = nb.get_motion_module(C, "Vanilla", kw)
                holders.append(h.to(torch.bfloat16).eval())
                xs.append(torch.randn(2, C, 8, side, side, device=dev, dtype=torch.bfloat16))
            ctx = torch.randn(2, 77, 768, device=dev, dtype=torch.bfloat16)

            def pair():
                for h, x in zip(holders, xs):
                    h.mm(h.sp(x, encoder_hidden_states=ctx).sample, None, None)
            a = timed(pair) / 2
            for h in holders:
                nb.patch(h, carry_stats=True)
            b1 = timed(pair) / 2
            for h in holders:
                nb.patch_spatial(h, carry_stats=True)
            b = timed(pair) / 2
        print(f"C={C:5d} {side:2d}x{side:<2d}  A {a:8.1f} us   B' {b1:8.1f} us   B {b:8.1f} us   fusion saves {b1 - b:6.1f} us per block, "
              f"motion-module emission costs {b1 - a:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
