#!/usr/bin/env python
"""Accuracy of the two fp32 modes against the committed golden fixtures (reference outputs) and, at BASELINE config 1 size, against
each other: NMM_F32X3 (3 bf16 tcgen05 MMAs per product, default for channels % 64 == 0) vs NMM_F32 (FMA-pipe GEMM, the checker).
Development / evidence tool: writes gpurun_out/fp32_modes.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers  # noqa: E402
from oracle import motion_oracle as mo  # noqa: E402


def run(cfg, params, x, fma):
    m = helpers.mirror_module(cfg, params, "cuda")
    if fma:
        m.__dict__["_nmm_fp32_fma"] = True
    with torch.no_grad():
        return m(x.cuda(), None, None).double().cpu()


lines = [f"{'case':28s} {'max|x3 - ref|':>14s} {'max|fma - ref|':>15s} {'max|x3 - fma|':>14s}   (bar: 1e-4 vs the fp32 reference)"]
for name in helpers.golden_names():
    fx, cfg, params, x = helpers.load_golden(name)
    ref = fx["out_ref_fp32"].double()
    y3, yf = run(cfg, params, x, False), run(cfg, params, x, True)
    lines.append(f"{name:28s} {(y3 - ref).abs().max().item():14.3e} {(yf - ref).abs().max().item():15.3e} {(y3 - yf).abs().max().item():14.3e}")
    print(lines[-1], flush=True)
cfg = mo.MotionConfig(320)
params = mo.make_params(cfg, 0)
for shape in ((1, 320, 8, 64, 64),):
    x = mo.make_input(shape, 0)
    y3, yf = run(cfg, params, x, False), run(cfg, params, x, True)
    lines.append(f"{'config1 ' + 'x'.join(map(str, shape)):28s} {'-':>14s} {'-':>15s} {(y3 - yf).abs().max().item():14.3e}")
    print(lines[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "fp32_modes.txt"), "w").write("\n".join(lines) + "\n")
