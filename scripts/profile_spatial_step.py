#!/usr/bin/env python
"""The 16 spatial Transformer3DModel calls of one UNet step (bench.py's `spatial_transformer` leg) bracketed by cudaProfilerStart/Stop,
for use under ncu (same recipe as scripts/profile_step.py):

  ncu --profile-from-start off --clock-control none --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum --log-file gpurun_out/spatial_step_metrics.csv \
      python scripts/profile_spatial_step.py
  python scripts/ncu_summarise.py gpurun_out/spatial_step_metrics.csv profiles/r2_ncu_spatial_step_summary.txt /dev/null
Numbers printed by a run under ncu are never bench values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    mods, xs = [], []
    with torch.no_grad():
        for C, side, n in [(320, 64, 5), (640, 32, 5), (1280, 16, 5), (1280, 8, 1)]:
            for _ in range(n):
                with torch.device(dev):
                    m = nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=C // 8, in_channels=C, cross_attention_dim=768,
                                              unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
                mods.append(m.to(torch.bfloat16).eval())
                xs.append(torch.randn(2, C, 8, side, side, device=dev, dtype=torch.bfloat16))
        ctx = torch.randn(2, 77, 768, device=dev, dtype=torch.bfloat16)
        for m, x in zip(mods, xs):
            m(x, encoder_hidden_states=ctx)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for m, x in zip(mods, xs):
            m(x, encoder_hidden_states=ctx)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled", len(mods), "calls")


if __name__ == "__main__":
    main()
