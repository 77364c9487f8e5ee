#!/usr/bin/env python
"""BASELINE.json configs[4]: motion-module shape sweep (roofline characterisation) -- 320/640/1280 channels x 8/16 frames x
32^2/64^2 latents, batch 1, one module call per shape, bf16 (tensor-core path) and optionally fp32 (parity mode).
Reports device time (CUDA events, L2 flushed between repetitions), achieved TFLOP/s against the measured bf16 peak and the
module's HBM floor (3 passes over the activation + parameters) against the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp32", action="store_true", help="also run the fp32 parity mode")
    ap.add_argument("--out", default="gpurun_out/shape_sweep.txt")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = [f"# B=1, one module call; peaks: bf16 {peak_tf:.0f} TFLOP/s (burst), HBM {peak_bw:.0f} GB/s ({'measured' if peaks else 'fallback'})",
             f"{'mode':6s} {'C':>5s} {'F':>3s} {'side':>4s} {'tokens':>7s} {'GFLOP':>8s} {'us':>9s} {'TFLOP/s':>8s} {'%peak':>6s} {'floor MB':>9s} {'floor us':>8s} {'x floor':>7s}"]
    # fp32 activations run in two modes: f32x3 = every Linear as 3 bf16 tcgen05 MMAs per product (the default), f32fma = FMA-pipe GEMM (checker)
    modes = [("bf16", torch.bfloat16, False)] + ([("f32x3", torch.float32, False), ("f32fma", torch.float32, True)] if a.fp32 else [])
    with torch.no_grad():
        for label, dt, fma in modes:
            for C in (320, 640, 1280):
                with torch.device(dev):
                    m = nb.get_motion_module(C, "Vanilla", kw).to(dt).eval()
                if fma:
                    m.__dict__["_nmm_fp32_fma"] = True
                for F in (8, 16):
                    for side in (32, 64):
                        x = torch.randn(1, F, C, side, side, device=dev, dtype=dt).permute(0, 2, 1, 3, 4)
                        for _ in range(2):
                            m(x, None, None)
                        ts = []
                        for _ in range(5):
                            flush.zero_()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(); m(x, None, None); e1.record()
                            torch.cuda.synchronize()
                            ts.append(e0.elapsed_time(e1) * 1e3)
                        us = sorted(ts)[len(ts) // 2]
                        N = F * side * side
                        fl = wl.module_flops(C, N, F)
                        floor = wl.module_min_bytes(C, N, 2 if dt == torch.bfloat16 else 4)
                        floor_us = floor / (peak_bw * 1e3)
                        lines.append(f"{label:6s} {C:5d} {F:3d} {side:4d} {N:7d} {fl / 1e9:8.1f} {us:9.1f} {fl / us / 1e6:8.1f} "
                                     f"{100 * fl / us / 1e6 / peak_tf:6.1f} {floor / 1e6:9.1f} {floor_us:8.1f} {us / floor_us:7.1f}")
                        print(lines[-1], flush=True)
                        del x
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
