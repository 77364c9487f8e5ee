"""GPU (-m gpu): the one-kernel C = 320 module (csrc/fused_module.cu) against the CPU oracle, stage by stage and end to end.

Stages are fp32 snapshots of the TMEM-resident residual stream (nmm_forward_stage): 0 = after proj_in (motion_module.py:145),
1 + i = after attention block i (:213-217), 1 + A = after the feed-forward (:219).  The oracle runs in fp64 on the same bf16-rounded
inputs and weights; the kernel rounds its GEMM operands (tokens, LayerNorm outputs, q|k|v, P, context, GEGLU activations) to bf16,
hence the stage tolerances below (a layout / indexing bug shows up as O(1) errors).
"""
import pytest
import torch

import neurons_b200 as nb
from neurons_b200 import lib as nlib
from neurons_b200 import ops
from oracle import motion_oracle as mo
from tests import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"

CASES = [  # (B, F, H, W, A, max_len, layout)
    (1, 8, 4, 4, 2, 24, "bcfhw"),          # one tile
    (1, 8, 8, 8, 2, 24, "bfchw"),          # 4 tiles, UNet-style view
    (2, 16, 8, 8, 2, 24, "bfchw"),         # 16 frames: 8 positions per tile
    (1, 16, 4, 6, 1, 32, "bcfhw"),         # SparseCtrl variant: one attention block, max_len 32
    (3, 8, 16, 16, 2, 24, "bfchw"),        # 48 tiles
]


def _setup(B, F, H, W, A, max_len, layout, seed=11):
    cfg = mo.MotionConfig(320, attn_blocks=A, max_len=max_len)
    params = {k: helpers.round_bf16(v) for k, v in mo.make_params(cfg, seed).items()}
    x = helpers.round_bf16(mo.make_input((B, 320, F, H, W), seed + 1, layout=layout))
    return cfg, params, x


def _pack(cfg, params):
    ncfg = nb.ModuleConfig(cfg.channels, cfg.heads, cfg.layers, cfg.attn_blocks, cfg.pos_enc, cfg.max_len)
    return ncfg, ops.pack_params(ncfg, {k: v.to(DEV) for k, v in params.items()}, torch.bfloat16, torch.device(DEV))


def _maxabs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


@pytest.mark.parametrize("B,F,H,W,A,max_len,layout", CASES)
def test_fused_module_stages_vs_oracle(B, F, H, W, A, max_len, layout):
    cfg, params, x = _setup(B, F, H, W, A, max_len, layout)
    st = mo.forward_token_order(params, x, cfg, torch.float64)
    ncfg, packed = _pack(cfg, params)
    xd = x.to(DEV, torch.bfloat16)
    assert xd.stride() == x.stride()
    refs = [st.h0] + st.h_attn + st.h_ff
    worst = []
    for stage, ref in enumerate(refs):
        y, snap = ops.forward_packed(xd, packed, ncfg, stage=stage)
        torch.cuda.synchronize()
        err = _maxabs(snap, ref)
        worst.append(err)
        # bf16 operands, fp32 accumulation: a few 1e-3 per stage at these magnitudes (|h| ~ 1-4)
        assert err <= 2e-2, f"stage {stage}: max-abs {err:.3e} (all so far: {worst})"
    assert _maxabs(y, st.out) <= helpers.TOL_BF16


@pytest.mark.parametrize("B,F,H,W,A,max_len,layout", CASES)
def test_fused_module_matches_multi_kernel_pipeline(B, F, H, W, A, max_len, layout):
    cfg, params, x = _setup(B, F, H, W, A, max_len, layout, seed=21)
    ncfg, packed = _pack(cfg, params)
    xd = x.to(DEV, torch.bfloat16)
    n0 = nb.launch_count()
    y_one = ops.forward_packed(xd, packed, ncfg)
    assert nb.launch_count() - n0 == 2
    with nlib.options({nlib.OPT_FUSED_MODULE: 0}):
        y_multi = ops.forward_packed(xd, packed, ncfg)
    ref = mo.forward_token_order(params, x, cfg, torch.float64).out
    assert _maxabs(y_one, ref) <= helpers.TOL_BF16
    assert _maxabs(y_multi, ref) <= helpers.TOL_BF16
    assert _maxabs(y_one, y_multi) <= 2 ** -6 * ref.abs().max().item()
    # deterministic: same bits on a second run, and with the output given in the caller's own (contiguous b c f h w) layout
    assert torch.equal(y_one, ops.forward_packed(xd, packed, ncfg))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=DEV)
    assert torch.equal(ops.forward_packed(xd, packed, ncfg, out=out), y_one)


def test_fused_module_ineligible_shapes_take_the_pipeline():
    # 24 frames / ragged positions: not tileable as (128 / F) positions x F frames -> multi-kernel path, same results bar
    for (B, F, H, W) in [(1, 24, 2, 2), (1, 8, 3, 5), (1, 16, 3, 3)]:
        cfg, params, x = _setup(B, F, H, W, 2, 24, "bcfhw", seed=31)
        ncfg, packed = _pack(cfg, params)
        n0 = nb.launch_count()
        y = ops.forward_packed(x.to(DEV, torch.bfloat16), packed, ncfg)
        assert nb.launch_count() - n0 > 2
        assert _maxabs(y, mo.forward_token_order(params, x, cfg, torch.float64).out) <= helpers.TOL_BF16
        with pytest.raises(nlib.NmmError):
            ops.forward_packed(x.to(DEV, torch.bfloat16), packed, ncfg, stage=0)
