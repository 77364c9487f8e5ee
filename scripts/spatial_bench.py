#!/usr/bin/env python
"""Spatial transformer per UNet level: in-stream time of one Transformer3DModel call (bf16, config-2 shapes: CFG batch 2,
8 frames, 64 x 64 latent), the per-kernel split from the library's CUDA-event profiler, and stock torch eager (the reference's op sequence,
cuBLAS + ATen, materialised scores) on the same GPU for the same call.  Development / profiles tool."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import lib as nlib  # noqa: E402
from oracle import spatial_oracle as so  # noqa: E402   (eager GPU baseline + flop count only; never on the product path)


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = []
    # calls per UNet step: 2 down + 3 up at each CrossAttn level, 1 in the mid block (unet.py:157-258)
    for C, side, n_calls in [(320, a.latent, 5), (640, a.latent // 2, 5), (1280, a.latent // 4, 5), (1280, a.latent // 8, 1)]:
        cfg = so.SpatialConfig(C, 8, 1, 768, True)
        P = side * side
        N = a.batch * a.frames * P
        with torch.no_grad():
            mods = []
            for i in range(2):
                m = nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=C // 8, in_channels=C, cross_attention_dim=768,
                                          unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
                mods.append(m.to(dev).to(torch.bfloat16).eval())
            xs = [torch.randn(a.batch, C, a.frames, side, side, device=dev, dtype=torch.bfloat16) for _ in range(2)]
            ctx = torch.randn(a.batch, 77, 768, device=dev, dtype=torch.bfloat16)

            def ours():
                for m, x in zip(mods, xs):
                    m(x, encoder_hidden_states=ctx)
            us = timed(ours, a.reps) / 2
            nlib.profile_begin()
            ours()
            prof = nlib.profile_end()
            fl = so.flops(cfg, a.batch, a.frames, P, 77)
            attn_fl = 4.0 * N * P * C + 4.0 * N * 77 * C
            row = dict(C=C, side=side, tokens=N, us_per_call=us, tflops=fl / us / 1e6, calls_per_step=n_calls, flops=fl, attention_flops=attn_fl,
                       kernels={k: dict(launches=v["launches"] // 2, us=v["total_ms"] * 1e3 / 2, tflops=(v["flops"] / 2) / max(v["total_ms"] * 1e-3 / 2, 1e-12) / 1e12)
                                for k, v in prof.items() if v["launches"]})
            if not a.no_eager:
                p = {k: v.detach() for k, v in mods[0].state_dict().items()}

                def eager():
                    so.forward_reference_order(p, xs[0], ctx, cfg)
                try:
                    row["eager_us_per_call"] = timed(eager, max(2, a.reps // 3))
                except torch.OutOfMemoryError:
                    row["eager_us_per_call"] = None
            out.append(row)
        sa = row["kernels"].get("spatial_attention", {})
        print(f"C={C:5d} side={side:3d} N={N:6d}  {us:8.1f} us/call  {row['tflops']:7.1f} TF/s  x{n_calls}  | attention kernels {sa.get('us', 0):8.1f} us "
              f"{sa.get('tflops', 0):6.1f} TF/s | eager {row.get('eager_us_per_call')}", flush=True)
        for k, v in row["kernels"].items():
            print(f"      {k:24s} x{v['launches']:2d} {v['us']:9.1f} us  {v['tflops']:7.1f} TF/s")
    step = sum(r["us_per_call"] * r["calls_per_step"] for r in out) / 1e3
    print(f"16 spatial-transformer calls of one UNet step: {step:.2f} ms, {sum(r['flops'] * r['calls_per_step'] for r in out) / step / 1e9:.1f} TF/s")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(dict(rows=out, step_ms=step), f, indent=1)


if __name__ == "__main__":
    main()
