#!/usr/bin/env python
"""Per-level module time inside a pipelined stream (PDL active, no per-kernel events): the in-situ cost of one motion-module
call at each UNet level of the bench workload, next to its algorithmic FLOPs.  Development tool."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ln-fold", action="store_true", help="fold the LayerNorms into the QKV / GEGLU GEMMs (multi-kernel pipeline levels)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = 0.0
    for C, side, n_calls in [(320, a.latent, 5), (640, a.latent // 2, 5), (1280, a.latent // 4, 5), (1280, a.latent // 8, 5)]:
        with torch.no_grad():
            with torch.device(dev):
                mods = [nb.get_motion_module(C, "Vanilla", kw).to(torch.bfloat16).eval() for _ in range(4)]
            if a.ln_fold:
                for m in mods:
                    m.__dict__["_nmm_ln_fold"] = True
            xs = [torch.randn(a.batch, a.frames, C, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4) for _ in range(4)]
            for m, x in zip(mods, xs):
                m(x, None, None)
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(a.reps):
                for m, x in zip(mods, xs):
                    m(x, None, None)
            e1.record()
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (a.reps * 4)
        fl = wl.module_flops(C, a.batch * a.frames * side * side, a.frames)
        total += us * n_calls
        print(f"C={C:5d} side={side:3d} M={a.batch * a.frames * side * side:6d}  {us:8.1f} us/call  {fl / us / 1e6:7.1f} TF/s   x{n_calls} = {us * n_calls / 1e3:.2f} ms/step", flush=True)
    print(f"sum over the 20 calls of a step: {total / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
