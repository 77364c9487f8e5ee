"""CPU (-m "not gpu"): neurons_b200.patch() on REFERENCE-BUILT models -- the unmodified UNet3DConditionModel and SparseControlNetModel
imported through oracle/unet_shim.py (build container only; skipped where /root/reference is absent, e.g. on the GPU box).

Covers the drop-in boundary of SURVEY.md 8(b): every VanillaTemporalModule the reference constructs (unet_blocks.py:244,358,470,603,718
via get_motion_module; sparse_controlnet.py) is found and rebound, its configuration is derived correctly (UNet: 2 attention blocks,
max_len 24 -- unet.py:157,183,236 + inference-v3.yaml; ControlNet: 1 block, max_len 32 -- sparsectrl/latent_condition.yaml:11-17), the module
tree / state_dict the reference's loaders walk stays untouched, and unsupported variants raise at patch time.
"""
import pytest
import torch

import neurons_b200 as nb
from oracle import ref_shim, unet_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def small_unet():
    # the full-width UNet (1.28 G parameters) takes ~20 s to construct on CPU; the structural checks only need the topology, so the widths
    # are divided by 4 here (same block types, same motion-module placement and kwargs); the full-size structure is checked by
    # oracle/gen_unet_golden.py when the fixtures are generated
    return unet_shim.build_unet(0, block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)


@pytest.fixture(scope="module")
def small_controlnet():
    return unet_shim.build_controlnet(0, block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)


def _motion(model):
    return [m for m in model.modules() if type(m).__name__ == "VanillaTemporalModule"]


def test_patch_finds_every_motion_module_of_the_reference_unet(small_unet):
    keys_before = list(small_unet.state_dict().keys())
    mods = _motion(small_unet)
    assert len(mods) == 20                                           # 2 per down block (4) + 3 per up block (4), none in the mid block
    assert nb.patch(small_unet) == 20
    assert list(small_unet.state_dict().keys()) == keys_before       # nothing added to / removed from the tree the checkpoint loaders walk
    chans = sorted({nb.config_of(m).channels for m in mods})
    assert chans == [64, 128, 256]
    for m in mods:
        cfg = nb.config_of(m)
        assert (cfg.heads, cfg.layers, cfg.attn_blocks, cfg.pos_enc, cfg.max_len) == (8, 1, 2, True, 24)
        assert m.forward.__func__.__name__ == "_fwd"                 # instance-level rebinding; the class is untouched
    assert type(mods[0]).forward.__name__ == "forward"
    # the walkers of unet.py:265-314 still see the attention sub-modules
    assert sum(1 for m in small_unet.modules() if hasattr(m, "set_attention_slice")) > 0


def test_patch_on_the_reference_controlnet(small_controlnet):
    mods = _motion(small_controlnet)
    assert len(mods) == 8
    assert nb.patch(small_controlnet) == 8
    for m in mods:
        cfg = nb.config_of(m)
        assert (cfg.heads, cfg.layers, cfg.attn_blocks, cfg.pos_enc, cfg.max_len) == (8, 1, 1, True, 32)


def test_patched_reference_module_has_no_cpu_path(small_unet):
    nb.patch(small_unet)
    m = _motion(small_unet)[0]
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 64, 8, 4, 4), None, None)
    with torch.no_grad(), pytest.raises(AssertionError, match="ndim=5"):      # same invariant / message as motion_module.py:135
        m(torch.zeros(1, 64, 8, 4), None, None)


def test_unsupported_reference_variants_raise_at_patch_time():
    ref = ref_shim.load_reference_motion_module()
    kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Cross"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=True)
    holder = torch.nn.ModuleList([ref.get_motion_module(64, "Vanilla", kw)])
    with pytest.raises(NotImplementedError):
        nb.patch(holder)
    with pytest.raises(ValueError):
        nb.get_motion_module(64, "NotVanilla", {})                   # same as the reference factory (motion_module.py:45)


def test_load_state_dict_into_patched_reference_module_repacks(small_unet):
    """The cache key carries tensor._version: load_state_dict after a warm-up (the reference's load_weights order) cannot serve stale
    packed weights.  (CPU-only check of the key; the numerical check is tests/test_gpu_parity.py::test_inplace_updates_...)"""
    from neurons_b200.motion_module import _Engine, _param_tensors
    m = _motion(small_unet)[0]
    t = _param_tensors(m)
    x = torch.zeros(1)
    k0 = _Engine._key(x, t)
    sd = {k: v.clone() + 0.5 for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert _Engine._key(x, _param_tensors(m)) != k0
