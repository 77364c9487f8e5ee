#!/usr/bin/env python
"""Time one C = 320 module call on the one-kernel path and on the multi-kernel pipeline (CUDA events, L2 flushed between calls)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import lib as nlib  # noqa: E402
from neurons_b200 import workloads as wl  # noqa: E402

dev = torch.device("cuda", 0)
kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
          temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, F, side) in [(2, 8, 64), (1, 8, 64), (2, 16, 32), (8, 16, 32), (2, 8, 16)]:
    with torch.no_grad():
        with torch.device(dev):
            m = nb.get_motion_module(320, "Vanilla", kw).to(torch.bfloat16).eval()
        x = torch.randn(B, F, 320, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4)
        for fused in (1, 0):
            with nlib.options({nlib.OPT_FUSED_MODULE: fused}):
                m(x, None, None)
                n0 = nb.launch_count()
                m(x, None, None)
                launches = nb.launch_count() - n0
                ts = []
                for r in range(10):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    m(x, None, None)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                fl = wl.module_flops(320, B * F * side * side, F)
                us = ts[len(ts) // 2]
                print(f"B={B} F={F} side={side} fused={fused}: {launches:2d} launches  median {us:8.1f} us  min {ts[0]:8.1f} us  {fl / us / 1e6:7.1f} TF/s", flush=True)
