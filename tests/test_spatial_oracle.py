"""CPU: the spatial-transformer oracle (oracle/spatial_oracle.py) against the committed reference fixtures and -- where the reference
tree is mounted -- against the live, unmodified reference class; the host mirror's checkpoint layout; patch_spatial on reference models."""
import glob
import os
import sys

import pytest
import torch

from oracle import spatial_oracle as so
from oracle import unet_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sp_golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "sp_*.pt")))
    assert names, "no spatial golden fixtures found"
    return names


def load_sp_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)
    m = fx["meta"]
    cfg = so.SpatialConfig(m["channels"], m["heads"], m["layers"], m["ctx_dim"], m["conv_proj"])
    params = so.make_params(cfg, m["param_seed"])
    x, ctx = so.make_inputs(cfg, m["batch"], m["frames"], m["height"], m["width"], m["ctx_len"], m["input_seed"], layout=m["layout"])
    pc = float(sum(v.double().abs().sum() for v in params.values()))
    xc = float(x.double().abs().sum()) + float(ctx.double().abs().sum())
    assert abs(pc - m["params_checksum"]) <= 1e-9 * max(1.0, abs(pc)), "seeded weight generator drifted from the fixture"
    assert abs(xc - m["input_checksum"]) <= 1e-9 * max(1.0, abs(xc)), "seeded input generator drifted from the fixture"
    return fx, cfg, params, x, ctx


@pytest.mark.parametrize("name", sp_golden_names())
def test_oracle_matches_reference_fixture(name):
    fx, cfg, params, x, ctx = load_sp_golden(name)
    with torch.no_grad():
        y = so.forward_reference_order(params, x, ctx, cfg)
    assert y.shape == fx["out_ref_fp32"].shape
    assert (y - fx["out_ref_fp32"]).abs().max().item() <= 2e-5          # same op sequence; ATen may pick another GEMM path per shape
    rb = lambda t: t.to(torch.bfloat16).float()
    with torch.no_grad():
        yb = so.forward_reference_order({k: rb(v) for k, v in params.items()}, rb(x), rb(ctx), cfg)
    assert (yb - fx["out_ref_bf16in"]).abs().max().item() <= 2e-5


@pytest.mark.skipif(not unet_shim.available(), reason="reference tree not mounted")
def test_pin_against_live_reference():
    from oracle import gen_spatial_golden as gg
    for C, F, H, W, B, Lc, layers, conv in ((320, 2, 4, 4, 2, 77, 1, True), (640, 1, 3, 5, 1, 5, 2, False)):
        cfg = so.SpatialConfig(C, 8, layers, 768, conv)
        params = so.make_params(cfg, 3)
        x, ctx = so.make_inputs(cfg, B, F, H, W, Lc, 4)
        with torch.no_grad():
            m = gg.build_reference(cfg)
            m.load_state_dict(params, strict=True)                      # the oracle's key list == the live class's state_dict
            ref = m(x, encoder_hidden_states=ctx).sample
            ours = so.forward_reference_order(params, x, ctx, cfg)
            assert ours.shape == ref.shape and ours.stride() == ref.stride()
            assert (ours - ref).abs().max().item() <= 2e-6
            m64 = gg.build_reference(cfg).double()
            m64.load_state_dict({k: v.double() for k, v in params.items()})
            ref64 = m64(x.double(), encoder_hidden_states=ctx.double()).sample
            ours64 = so.forward_reference_order({k: v.double() for k, v in params.items()}, x.double(), ctx.double(), cfg)
            assert (ours64 - ref64).abs().max().item() <= 1e-12


def test_mirror_checkpoint_layout():
    import neurons_b200 as nb
    for conv in (True, False):
        cfg = so.SpatialConfig(320, 8, 2, 768, conv)
        m = nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, num_layers=2, cross_attention_dim=768,
                                  use_linear_projection=not conv, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
        sd = m.state_dict()
        shapes = so.param_shapes(cfg)
        assert set(sd.keys()) == set(shapes.keys())
        for k, shp in shapes.items():
            assert tuple(sd[k].shape) == shp, k
        m.load_state_dict(so.make_params(cfg, 1), strict=True)
        assert nb.spatial_config_of(m) == nb.SpatialConfig(320, 8, 2, 768)
    with pytest.raises(NotImplementedError):
        nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, cross_attention_dim=768,
                              unet_use_cross_frame_attention=True, unet_use_temporal_attention=False)
    x = torch.zeros(1, 320, 1, 2, 2)
    with pytest.raises(RuntimeError):           # no CPU path
        m(x, encoder_hidden_states=torch.zeros(1, 77, 768))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 320, 2, 2), encoder_hidden_states=torch.zeros(1, 77, 768))


@pytest.mark.skipif(not unet_shim.available(), reason="reference tree not mounted")
def test_patch_spatial_on_reference_models():
    import neurons_b200 as nb
    small = dict(block_out_channels=(320, 640, 1280, 1280))
    with torch.no_grad():
        unet = unet_shim.build_unet(0, **small)
        assert nb.patch_spatial(unet) == 16                      # 2+2+2 down, 1 mid, 3+3+3 up (unet.py:157-258)
        cn = unet_shim.build_controlnet(0)
        n_cn = sum(1 for m in cn.modules() if type(m).__name__ == "Transformer3DModel")
        assert nb.patch_spatial(cn) == n_cn == 7           # 2+2+2 down, 1 mid (sparse_controlnet.py)
    tr = next(m for m in unet.modules() if type(m).__name__ == "Transformer3DModel")
    assert nb.spatial_config_of(tr).ctx_dim == 768
    with pytest.raises(RuntimeError):           # patched forward has no CPU path: it must raise, not fall back
        tr(torch.zeros(1, tr.norm.num_channels, 1, 2, 2), encoder_hidden_states=torch.zeros(1, 77, 768))
    att = sys.modules["animatediff.models.attention"]
    bad = att.Transformer3DModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, cross_attention_dim=768,
                                 unet_use_cross_frame_attention=False, unet_use_temporal_attention=True)
    with pytest.raises(NotImplementedError):
        nb.patch_spatial(bad)
