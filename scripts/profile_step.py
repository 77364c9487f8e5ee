#!/usr/bin/env python
"""One motion-module UNet step (the bench.py workload) bracketed by cudaProfilerStart/Stop, for use under ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:linear_tc -c 12 -o gpurun_out/prof_gemm \
      python scripts/profile_step.py --calls 2
Numbers printed by a run under ncu are never bench values.
"""
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
    ap.add_argument("--calls", type=int, default=20, help="how many of the step's 20 calls to run")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--only", default="", help="comma-separated indices (into the first --calls calls) to profile; default all")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    kwargs = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                  temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1,
                  zero_initialize=False)
    torch.manual_seed(0)
    calls = wl.unet_step_calls(a.latent)[: a.calls]
    if a.only:
        calls = [calls[int(i)] for i in a.only.split(",")]
    mods, xs = [], []
    with torch.no_grad():
        for c in calls:
            with torch.device(dev):
                m = nb.get_motion_module(c.channels, "Vanilla", kwargs)
            mods.append(m.to(torch.bfloat16).eval())
            xs.append(torch.randn(2, 8, c.channels, c.side, c.side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4))
        for m, x in zip(mods, xs):      # warm-up: packs parameters, sets kernel attributes
            m(x, None, None)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for m, x in zip(mods, xs):
            m(x, None, None)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled", len(calls), "calls")


if __name__ == "__main__":
    main()
