#!/usr/bin/env python
"""SURVEY 8(f) row N1, timed: GroupNorm statistics carried from the motion module to the next InflatedGroupNorm (ResnetBlock3D.norm1)
instead of recomputed.  Per UNet level of the bench workload (bf16, CFG batch 2, 8 frames):
    module              one motion-module call                        (gn_stats over x + the module)
    module + y sums     ... that also emits the sums of y             (epilogue emission + reduction kernel)
    module, x sums in   ... with the sums of x handed in              (no statistics pass over x)
    norm1               InflatedGroupNorm + SiLU of y                 (statistics pass + apply pass)
    norm1, carried      ... with the carried sums                     (apply pass only)
Each variant is captured in a CUDA graph over 4 distinct buffer sets and replayed (device time per call, no host overhead)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import ops  # noqa: E402


def timed(fns, reps=5):
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for fn in fns:
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


def main():
    dev = torch.device("cuda", 0)
    kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
    lines = [f"{'level':18s} {'module':>9s} {'+ y sums':>9s} {'x sums in':>10s} | {'norm1':>8s} {'carried':>8s} | pair before -> after (us)"]
    with torch.no_grad():
        for C, side in ((320, 64), (640, 32), (1280, 16), (1280, 8)):
            B, F = 2, 8
            with torch.device(dev):
                m = nb.get_motion_module(C, "Vanilla", kw).to(torch.bfloat16).eval()
            xs = [torch.randn(B, F, C, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4) for _ in range(4)]
            m(xs[0], None, None)
            eng = m.__dict__["_nmm_engine"]
            w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            ys = [m(x, None, None) for x in xs]
            ysums = [torch.empty(B * F * 32, 2, dtype=torch.float64, device=dev) for _ in xs]
            xsums = [ops.groupnorm_sums(x) for x in xs]
            for x, s in zip(xs, ysums):
                ops.forward_packed(x, eng.packed, eng.cfg, y_sums=s)
            outs = [torch.empty(y.shape, dtype=y.dtype, device=dev) for y in ys]
            t_mod = timed([lambda x=x: ops.forward_packed(x, eng.packed, eng.cfg) for x in xs])
            t_emit = timed([lambda x=x, s=s: ops.forward_packed(x, eng.packed, eng.cfg, y_sums=s) for x, s in zip(xs, ysums)])
            t_xin = timed([lambda x=x, s=s: ops.forward_packed(x, eng.packed, eng.cfg, x_sums=s) for x, s in zip(xs, xsums)])
            t_norm = timed([lambda y=y, o=o: ops.inflated_groupnorm(y, w, b, 1e-5, silu=True, out=o) for y, o in zip(ys, outs)])
            t_carr = timed([lambda y=y, o=o, s=s: ops.inflated_groupnorm(y, w, b, 1e-5, silu=True, out=o, sums=s) for y, o, s in zip(ys, outs, ysums)])
            lines.append(f"C={C:5d} {side:3d}x{side:<3d}   {t_mod:9.1f} {t_emit:9.1f} {t_xin:10.1f} | {t_norm:8.1f} {t_carr:8.1f} | {t_mod + t_norm:8.1f} -> {t_emit + t_carr:8.1f}")
            print(lines[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "n1_bench.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
