#!/usr/bin/env python
"""temporal attention + blend of the blurry-video decoder blocks (t = 6) -- nmm_decoder_temporal_attention against the
reference's op sequence (restated oracle) under stock torch eager on the same GPU, at the decoder's three widths.  Development / profiles tool."""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurons_b200 import video_decoder as vd  # noqa: E402
from oracle import decoder_oracle as do  # noqa: E402   (eager anchor only)
from tests.test_video_decoder import FakeAttention  # noqa: E402


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
    for dtype in (torch.bfloat16, torch.float32):
        for C, side, b in ((128, 28, 2), (64, 56, 2), (32, 112, 2)):
            cfg = do.DecoderAttnConfig(C)
            params = do.make_params(cfg, 1)
            attn = FakeAttention(cfg, params).cuda().to(dtype).eval()
            w = nn.Parameter(torch.tensor([0.5])).cuda()
            x = torch.randn(b * 6, C, side, side, device="cuda", dtype=dtype)
            pd = {k: v.cuda().to(dtype) for k, v in params.items()}
            with torch.no_grad():
                ours = timed(lambda: vd.temporal_attention_blend(x, attn, w, 6))
                eager = timed(lambda: do.temporal_blend_reference_order(pd, x, 0.5, 6, cfg))
            mb = 2 * x.numel() * x.element_size() / 1e6
            print(f"{str(dtype)[6:]:9s} C={C:4d} {side:3d}x{side:<3d} b={b} t=6  ours {ours:7.1f} us ({mb / ours * 1e3:6.1f} GB/s of x + y)   eager {eager:7.1f} us   x{eager / ours:4.1f}",
                  flush=True)


if __name__ == "__main__":
    main()
