"""CPU: the oracle against the committed golden fixtures (made from the unmodified reference, oracle/gen_golden.py)
and -- when the reference tree is present (build container only) -- against the reference itself."""
import pytest
import torch

from oracle import motion_oracle as mo
from oracle import ref_shim
from tests import helpers


@pytest.mark.parametrize("name", helpers.golden_names())
def test_reference_order_matches_golden_fp32(name):
    fx, cfg, params, x = helpers.load_golden(name)
    with torch.no_grad():
        y = mo.forward_reference_order(params, x, cfg)
    ref = fx["out_ref_fp32"]
    assert y.shape == ref.shape
    # same ops in the same order: bit-identical except where ATen picks a different GEMM path (<= 2 ulp-ish)
    assert (y - ref).abs().max().item() <= 2e-6


@pytest.mark.parametrize("name", helpers.golden_names())
def test_token_order_matches_golden(name):
    fx, cfg, params, x = helpers.load_golden(name)
    with torch.no_grad():
        st = mo.forward_token_order(params, x, cfg, torch.float64)
    if "out_ref_fp64" in fx:
        assert (st.out - fx["out_ref_fp64"]).abs().max().item() <= 1e-12
    assert (st.out - fx["out_ref_fp32"].double()).abs().max().item() <= 1e-5
    # bf16-rounded inputs / weights: the bf16-mode target
    pb = {k: helpers.round_bf16(v) for k, v in params.items()}
    with torch.no_grad():
        stb = mo.forward_token_order(pb, helpers.round_bf16(x), cfg, torch.float64)
    assert (stb.out - fx["out_ref_bf16in"].double()).abs().max().item() <= 1e-5


def test_output_is_bfchw_storage_view():
    # the reference returns a permuted view over [B,F,C,H,W] storage (motion_module.py:153-156)
    cfg = mo.MotionConfig(32, 8, 1, 1, True, 24)
    x = mo.make_input((2, 32, 3, 2, 2), 1)
    y = mo.forward_reference_order(mo.make_params(cfg, 1), x, cfg)
    assert y.shape == x.shape
    assert y.permute(0, 2, 1, 3, 4).is_contiguous()


def test_layout_invariance():
    cfg = mo.MotionConfig(64, 8, 1, 2, True, 24)
    p = mo.make_params(cfg, 2)
    a = mo.forward_reference_order(p, mo.make_input((1, 64, 4, 3, 3), 9, layout="bcfhw"), cfg)
    b = mo.forward_reference_order(p, mo.make_input((1, 64, 4, 3, 3), 9, layout="bfchw"), cfg)
    assert (a - b).abs().max().item() <= 2e-6


def test_invariants_raise():
    cfg = mo.MotionConfig(32, 8, 1, 1, True, 4)
    p = mo.make_params(cfg, 0)
    with pytest.raises(AssertionError):
        mo.forward_reference_order(p, torch.zeros(1, 32, 2, 2), cfg)            # ndim != 5, motion_module.py:135
    with pytest.raises(ValueError):
        mo.forward_reference_order(p, torch.zeros(1, 32, 5, 2, 2), cfg)         # frames > max_len


def test_flops_formula():
    # config 1 = 148.3 GFLOP
    assert abs(mo.flops(mo.MotionConfig(320), 1, 8, 64, 64) / 1e9 - 148.3) < 0.1


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("C,F,side,A,L,ml,layout", [(64, 8, 4, 2, 1, 24, "bcfhw"), (96, 16, 2, 1, 2, 32, "bfchw")])
def test_pin_against_live_reference(C, F, side, A, L, ml, layout):
    from oracle.gen_golden import build_reference
    cfg = mo.MotionConfig(C, 8, L, A, True, ml)
    params = mo.make_params(cfg, 21)
    x = mo.make_input((2, C, F, side, side), 22, layout=layout)
    with torch.no_grad():
        m = build_reference(cfg)
        missing, unexpected = m.load_state_dict(params, strict=False)
        assert not missing and not unexpected
        assert list(m.state_dict().keys()) == list(params.keys())
        y_ref = m(x, None, None)
        y = mo.forward_reference_order(params, x, cfg)
        assert y.stride() == y_ref.stride()
        assert (y - y_ref).abs().max().item() <= 2e-6
        m64 = build_reference(cfg).double()
        m64.load_state_dict({k: v.double() for k, v in params.items()}, strict=False)
        st = mo.forward_token_order(params, x, cfg, torch.float64)
        assert (st.out - m64(x.double(), None, None)).abs().max().item() <= 1e-12
