#!/usr/bin/env python
"""How fast can the e2e arm's host<->device traffic go on its own?  Copies the bench step's 20 inputs H2D and 20 outputs D2H
(380 MB each way, pinned memory) on two streams, alone and concurrently.  Development probe for DESIGN.md section 7."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from neurons_b200 import workloads as wl

dev = torch.device("cuda", 0)
calls = wl.unet_step_calls(64)
hs = [torch.empty((2, 8, c.channels, c.side, c.side), dtype=torch.bfloat16).pin_memory() for c in calls]
ho = [torch.empty_like(h).pin_memory() for h in hs]
ds = [torch.empty_like(h, device=dev) for h in hs]
nbytes = sum(h.numel() * 2 for h in hs)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(len(hs)):
            if h2d:
                with torch.cuda.stream(s1):
                    ds[i].copy_(hs[i], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    ho[i].copy_(ds[i], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

for name, a, b in [("H2D only", True, False), ("D2H only", False, True), ("both directions", True, True)]:
    run(a, b, 2)
    t = run(a, b)
    print(f"{name:16s} {t * 1e3:7.2f} ms per step-worth ({nbytes / 1e6:.0f} MB each way) -> {nbytes / t / 1e9:6.1f} GB/s per direction")
