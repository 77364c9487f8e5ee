"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/sp_*.pt from the UNMODIFIED reference Transformer3DModel
(/root/reference/animatediff/models/attention.py, imported through oracle/unet_shim.py).

Run in the build container (needs /root/reference):   python -m oracle.gen_spatial_golden
Each fixture: meta (config / shape / seeds / checksums of the regenerated weights and inputs), out_ref_fp32 (reference fp32),
out_ref_bf16in (reference in fp32 on bf16-ROUNDED weights and inputs: the bf16-mode bar's comparand).  Weights and inputs are not
stored: oracle.spatial_oracle.make_params / make_inputs regenerate them from the seeds.
"""
from __future__ import annotations

import os
import sys

import torch

from . import spatial_oracle as so
from . import unet_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (C, heads, F, H, W, B, ctx_len, layers, layout, conv_proj)
CASES = {
    "sp_c320_f2_8x8":        (320, 8, 2, 8, 8, 2, 77, 1, "bcfhw", True),     # d_h 40, P = 64: one key tile, GroupNorm fused into proj_in
    "sp_c320_f1_16x12":      (320, 8, 1, 16, 12, 1, 77, 1, "bcfhw", True),   # P = 192: 3 key tiles, a ragged 128-query tile
    "sp_c640_f2_5x4_view":   (640, 8, 2, 5, 4, 2, 77, 1, "bfchw", True),     # d_h 80, P = 20 (ragged everything), [B,F,C,H,W]-storage view input
    "sp_c1280_f1_4x4":       (1280, 8, 1, 4, 4, 2, 77, 1, "bcfhw", True),    # d_h 160, P = 16
    "sp_c320_f2_4x4_l2_lin": (320, 8, 2, 4, 4, 1, 13, 2, "bcfhw", False),    # two blocks, use_linear_projection, short context
}
PARAM_SEED, INPUT_SEED = 23, 9


def build_reference(cfg: so.SpatialConfig):
    unet_shim.load()
    att = sys.modules["animatediff.models.attention"]
    return att.Transformer3DModel(num_attention_heads=cfg.heads, attention_head_dim=cfg.head_dim, in_channels=cfg.channels, num_layers=cfg.layers,
                                  cross_attention_dim=cfg.ctx_dim, use_linear_projection=not cfg.conv_proj, unet_use_cross_frame_attention=False,
                                  unet_use_temporal_attention=False).eval()


def checksum(t) -> float:
    return float(t.double().abs().sum())


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def generate(name: str):
    C, heads, F, H, W, B, Lc, layers, layout, conv = CASES[name]
    cfg = so.SpatialConfig(C, heads, layers, 768, conv)
    params = so.make_params(cfg, PARAM_SEED)
    x, ctx = so.make_inputs(cfg, B, F, H, W, Lc, INPUT_SEED, layout=layout)
    with torch.no_grad():
        m = build_reference(cfg)
        m.load_state_dict(params, strict=True)
        out32 = m(x, encoder_hidden_states=ctx).sample.contiguous()
        mb = build_reference(cfg)
        mb.load_state_dict({k: round_bf16(v) for k, v in params.items()}, strict=True)
        outb = mb(round_bf16(x), encoder_hidden_states=round_bf16(ctx)).sample.contiguous()
    fx = {"meta": dict(name=name, channels=C, heads=heads, frames=F, height=H, width=W, batch=B, ctx_len=Lc, layers=layers, layout=layout,
                       conv_proj=conv, ctx_dim=768, param_seed=PARAM_SEED, input_seed=INPUT_SEED,
                       params_checksum=float(sum(checksum(v) for v in params.values())), input_checksum=checksum(x) + checksum(ctx),
                       torch_version=torch.__version__),
          "out_ref_fp32": out32, "out_ref_bf16in": outb}
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.save(fx, os.path.join(GOLDEN_DIR, name + ".pt"))
    return fx


def main():
    if not unet_shim.available():
        print("reference tree not available; cannot generate goldens", file=sys.stderr)
        sys.exit(1)
    for name in CASES:
        fx = generate(name)
        print(f"{name}: out {tuple(fx['out_ref_fp32'].shape)} absmax {fx['out_ref_fp32'].abs().max():.4f} "
              f"bf16in-vs-fp32 {float((fx['out_ref_fp32'] - fx['out_ref_bf16in']).abs().max()):.4f}")


if __name__ == "__main__":
    main()
