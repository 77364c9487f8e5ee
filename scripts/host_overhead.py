#!/usr/bin/env python
"""Host enqueue time vs device time of one module call per UNet level (is a level launch-bound?)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402

dev = torch.device("cuda", 0)
kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
          temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
for C, side in [(320, 64), (640, 32), (1280, 16), (1280, 8)]:
    with torch.no_grad():
        with torch.device(dev):
            m = nb.get_motion_module(C, "Vanilla", kw).to(torch.bfloat16).eval()
        x = torch.randn(2, 8, C, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4)
        for _ in range(3):
            m(x, None, None)
        torch.cuda.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            m(x, None, None)
        e1.record()
        t_host = (time.perf_counter() - t0) / reps * 1e6
        torch.cuda.synchronize()
        t_dev = e0.elapsed_time(e1) / reps * 1e3
        # the same through a CUDA graph
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(10):
                    m(x, None, None)
        torch.cuda.current_stream().wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t_graph = e0.elapsed_time(e1) / 50 * 1e3
    print(f"C={C:5d} side={side:3d}: host enqueue {t_host:7.1f} us/call   stream time {t_dev:7.1f} us/call   graph replay {t_graph:7.1f} us/call", flush=True)
