"""Shared helpers for the parity tests (the oracle is the checker; see oracle/motion_oracle.py header)."""
import glob
import os

import torch

from oracle import motion_oracle as mo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_FP32 = 1e-4     # north_star: max-abs vs the fp32 reference in fp32 mode
TOL_BF16 = 2e-2     # north_star: max-abs vs the reference on identical (bf16-rounded) inputs and weights


def golden_names():
    """Module-level fixtures (oracle/gen_golden.py).  `unet_step_inputs.pt` has another schema (oracle/gen_unet_golden.py,
    tests/test_unet_activations.py) and is not one of them."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "c*_f*.pt")))
    assert names, "no golden fixtures found"
    return names


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)
    m = fx["meta"]
    cfg = mo.MotionConfig(m["channels"], m["heads"], m["layers"], m["attn_blocks"], True, m["max_len"])
    params = mo.make_params(cfg, m["param_seed"])
    x = mo.make_input((m["batch"], m["channels"], m["frames"], m["height"], m["width"]), m["input_seed"], layout=m["layout"])
    # the seeded generators must reproduce what the fixture was made from
    pc = float(sum(v.double().abs().sum() for v in params.values()))
    xc = float(x.double().abs().sum())
    assert abs(pc - m["params_checksum"]) <= 1e-9 * max(1.0, abs(pc)), "seeded weight generator drifted from the fixture"
    assert abs(xc - m["input_checksum"]) <= 1e-9 * max(1.0, abs(xc)), "seeded input generator drifted from the fixture"
    return fx, cfg, params, x


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def mirror_module(cfg, params, device=None, dtype=torch.float32):
    """neurons_b200.VanillaTemporalModule carrying `params` (state_dict load = checkpoint-layout check)."""
    import neurons_b200 as nb
    m = nb.get_motion_module(cfg.channels, "Vanilla", dict(
        num_attention_heads=cfg.heads, num_transformer_block=cfg.layers,
        attention_block_types=("Temporal_Self",) * cfg.attn_blocks, cross_frame_attention_mode=None,
        temporal_position_encoding=cfg.pos_enc, temporal_position_encoding_max_len=cfg.max_len,
        temporal_attention_dim_div=1, zero_initialize=False))
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    m = m.eval()
    if device is not None:
        m = m.to(device)
    if dtype != torch.float32:
        m = m.to(dtype)
    return m
