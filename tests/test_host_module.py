"""CPU: the host-side mirror of the reference interface (names, checkpoint layout, error behaviour)."""
import pytest
import torch

import neurons_b200 as nb
from neurons_b200 import motion_module as mm
from oracle import motion_oracle as mo
from tests import helpers

V3_KWARGS = dict(   # /root/reference/configs/inference/inference-v3.yaml:8-14
    num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
    temporal_position_encoding=True, temporal_attention_dim_div=1, zero_initialize=True)


def test_factory_and_errors():
    m = nb.get_motion_module(320, "Vanilla", V3_KWARGS)
    assert isinstance(m, nb.VanillaTemporalModule)
    with pytest.raises(ValueError):
        nb.get_motion_module(320, "Other", V3_KWARGS)               # motion_module.py:45
    with pytest.raises(AssertionError):
        nb.get_motion_module(64, "Vanilla", dict(V3_KWARGS, attention_block_types=["Spatial_Self"]))    # :256
    with pytest.raises(NotImplementedError):
        nb.get_motion_module(64, "Vanilla", dict(V3_KWARGS, attention_block_types=["Temporal_Cross"]))
    with pytest.raises(NotImplementedError):
        nb.get_motion_module(64, "Vanilla", dict(V3_KWARGS, temporal_attention_dim_div=2))


@pytest.mark.parametrize("C,A,L,ml", [(320, 2, 1, 24), (640, 1, 1, 32), (64, 2, 2, 24)])
def test_checkpoint_layout(C, A, L, ml):
    cfg = mo.MotionConfig(C, 8, L, A, True, ml)
    m = helpers.mirror_module(cfg, mo.make_params(cfg, 3))
    sd = m.state_dict()
    shapes = mo.param_shapes(cfg)          # listed from the live reference
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    # pos_encoder.pe is a non-persistent buffer (motion_module.py:239; stripped by util.py:116)
    assert not any("pos_encoder.pe" in k for k in sd)
    pe = m.temporal_transformer.transformer_blocks[0].attention_blocks[0].pos_encoder.pe
    assert tuple(pe.shape) == (1, ml, C)
    assert torch.equal(pe[0], mo.positional_encoding(ml, C))
    assert mm.config_of(m) == nb.ModuleConfig(C, 8, L, A, True, ml)


def test_param_count_matches_survey():
    m = nb.get_motion_module(320, "Vanilla", V3_KWARGS)
    assert sum(p.numel() for p in m.parameters()) == 2_259_520       # 2.26 M @ C=320


def test_zero_initialize():
    m = nb.get_motion_module(64, "Vanilla", V3_KWARGS)
    assert float(m.temporal_transformer.proj_out.weight.abs().sum()) == 0.0
    assert float(m.temporal_transformer.proj_out.bias.abs().sum()) == 0.0
    m2 = nb.get_motion_module(64, "Vanilla", dict(V3_KWARGS, zero_initialize=False))
    assert float(m2.temporal_transformer.proj_out.weight.abs().sum()) > 0.0


def test_forward_refuses_cpu_and_bad_rank():
    m = nb.get_motion_module(64, "Vanilla", V3_KWARGS).eval()
    with torch.no_grad():
        with pytest.raises(AssertionError):
            m(torch.zeros(1, 64, 4, 4), None, None)                  # ndim != 5 (:135)
        with pytest.raises(RuntimeError, match="CUDA"):
            m(torch.zeros(1, 64, 4, 2, 2), None, None)               # no CPU path


def test_patch_rebinds_only_motion_modules():
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.motion_modules = torch.nn.ModuleList([nb.get_motion_module(32, "Vanilla", V3_KWARGS) for _ in range(2)])
            self.other = torch.nn.Linear(4, 4)
    blk = Block()
    keys_before = list(blk.state_dict().keys())
    assert nb.patch(blk) == 2
    assert list(blk.state_dict().keys()) == keys_before            # nothing added to the module tree
    assert nb.invalidate(blk) == 0                                   # nothing packed yet


def test_set_attention_slice_walker_compat():
    # unet.set_attention_slice walks modules exposing set_attention_slice / sliceable_head_dim (unet.py:265-314)
    m = nb.get_motion_module(64, "Vanilla", V3_KWARGS)
    attn = m.temporal_transformer.transformer_blocks[0].attention_blocks[0]
    assert attn.sliceable_head_dim == 8
    attn.set_attention_slice(4)
    with pytest.raises(ValueError):
        attn.set_attention_slice(9)


def test_inflated_groupnorm_mirror_keeps_checkpoint_layout_and_has_no_cpu_path():
    import neurons_b200 as nb
    ref = torch.nn.GroupNorm(32, 64, eps=1e-5)
    ours = nb.InflatedGroupNorm(32, 64, eps=1e-5)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    with pytest.raises(RuntimeError, match="CUDA"):
        ours(torch.zeros(1, 64, 2, 2, 2))
    with pytest.raises(NotImplementedError):
        nb.InflatedGroupNorm(16, 64)

    class InflatedGroupNorm(torch.nn.GroupNorm):          # stands for the reference class (matched by name)
        pass
    model = torch.nn.Sequential(InflatedGroupNorm(32, 64), torch.nn.GroupNorm(32, 64), InflatedGroupNorm(8, 64))
    assert nb.patch_group_norms(model) == 1
