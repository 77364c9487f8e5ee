"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt from the UNMODIFIED reference module.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The fixtures travel to the GPU box, where the reference tree does not exist.

Each fixture holds, for one (config, shape, layout, seed):
  meta            config/shape/seeds + float64 checksums of the generated weights and input
                  (so a drift in the seeded generators is detected instead of silently mis-comparing)
  out_ref_fp32    reference VanillaTemporalModule (fp32) on the fp32 input            -> fp32-mode bar 1e-4
  out_ref_bf16in  reference module in fp32 fed bf16-ROUNDED weights and input         -> bf16-mode bar 2e-2
                  (north_star: "identical inputs and weights"; `pe` stays fp32 -- SURVEY 8(c))
  out_ref_fp64    reference module .double() on the fp32 input (small cases only)     -> oracle fp64 pin
Weights and inputs are NOT stored: oracle.motion_oracle.make_params / make_input regenerate them.
"""
from __future__ import annotations

import os
import sys

import torch

from . import motion_oracle as mo
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (C, F, H, W, B, attn_blocks, layers, max_len, layout, store_fp64)
CASES = {
    "c320_f8_8x8_a2":        (320, 8, 8, 8, 1, 2, 1, 24, "bcfhw", False),   # config-1 topology, small latent
    "c64_f16_4x4_a1_view":   (64, 16, 4, 4, 2, 1, 1, 32, "bfchw", True),    # SparseCtrl variant (A=1, max_len 32), UNet-style view
    "c640_f16_4x4_a2_l2":    (640, 16, 4, 4, 1, 2, 2, 24, "bfchw", False),  # two transformer blocks (class default)
    "c1280_f8_2x2_a1":       (1280, 8, 2, 2, 2, 1, 1, 32, "bcfhw", True),   # deepest level, P=4 (8-byte rows)
    "c320_f24_3x5_a2":       (320, 24, 3, 5, 1, 2, 1, 24, "bcfhw", False),  # f == max_len, ragged/odd spatial size
    "c32_f1_1x2_a2":         (32, 1, 1, 2, 1, 2, 1, 24, "bcfhw", True),     # degenerate: 1 frame, 2 positions, d_h = 4, 2-value groups
    "c320_f16_8x8_a2_view":  (320, 16, 8, 8, 1, 2, 1, 24, "bfchw", False),   # NEURONS shape class (16 frames), UNet-style view input
}
PARAM_SEED, INPUT_SEED = 11, 5


def build_reference(cfg: mo.MotionConfig):
    ref = ref_shim.load_reference_motion_module()
    kwargs = dict(
        num_attention_heads=cfg.heads,
        num_transformer_block=cfg.layers,
        attention_block_types=("Temporal_Self",) * cfg.attn_blocks,
        cross_frame_attention_mode=None,
        temporal_position_encoding=cfg.pos_enc,
        temporal_position_encoding_max_len=cfg.max_len,
        temporal_attention_dim_div=1,
        zero_initialize=False,
    )
    return ref.get_motion_module(cfg.channels, "Vanilla", kwargs).eval()


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def params_checksum(p) -> float:
    return float(sum(checksum(v) for v in p.values()))


def round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def generate(name: str):
    C, F, H, W, B, A, L, max_len, layout, store64 = CASES[name]
    cfg = mo.MotionConfig(C, 8, L, A, True, max_len)
    params = mo.make_params(cfg, PARAM_SEED)
    x = mo.make_input((B, C, F, H, W), INPUT_SEED, layout=layout)
    with torch.no_grad():
        m = build_reference(cfg)
        missing, unexpected = m.load_state_dict(params, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        out32 = m(x, None, None).contiguous()
        mb = build_reference(cfg)
        mb.load_state_dict({k: round_bf16(v) for k, v in params.items()}, strict=False)
        outb = mb(round_bf16(x), None, None).contiguous()
        fx = {
            "meta": dict(name=name, channels=C, frames=F, height=H, width=W, batch=B, attn_blocks=A, layers=L,
                         max_len=max_len, heads=8, layout=layout, param_seed=PARAM_SEED, input_seed=INPUT_SEED,
                         params_checksum=params_checksum(params), input_checksum=checksum(x),
                         torch_version=torch.__version__),
            "out_ref_fp32": out32,
            "out_ref_bf16in": outb,
        }
        if store64:
            m64 = build_reference(cfg).double()
            m64.load_state_dict({k: v.double() for k, v in params.items()}, strict=False)
            fx["out_ref_fp64"] = m64(x.double(), None, None).contiguous()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.save(fx, os.path.join(GOLDEN_DIR, name + ".pt"))
    return fx


def main():
    if not ref_shim.available():
        print("reference tree not available; cannot generate goldens", file=sys.stderr)
        sys.exit(1)
    torch.manual_seed(0)
    for name in CASES:
        fx = generate(name)
        print(f"{name}: out {tuple(fx['out_ref_fp32'].shape)} absmax {fx['out_ref_fp32'].abs().max():.4f} "
              f"bf16in-vs-fp32 {float((fx['out_ref_fp32'] - fx['out_ref_bf16in']).abs().max()):.4f}")


if __name__ == "__main__":
    main()
