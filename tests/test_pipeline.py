"""CPU: the stage-5 harness around the hot path (neurons_b200/pipeline.py): frame interpolation and clip indexing pinned
to the reference's own functions (compiled from the reference script where it lies), the stage-3 file formats, and the SparseCtrl
ControlNet residual plumbing driven through the UNMODIFIED reference UNet3DConditionModel / SparseControlNetModel (oracle/unet_shim.py)."""
import ast
import os

import pytest
import torch

from neurons_b200 import pipeline, sampler, sharding
from oracle import ref_shim, unet_shim


def _reference_functions(*names):
    """Functions of scripts/neuroclips_video_enhance.py compiled from the reference file (the script itself needs packages this image lacks)."""
    root = ref_shim.reference_root()
    path = os.path.join(root or "", "scripts", "neuroclips_video_enhance.py")
    if root is None or not os.path.isfile(path):
        return None
    tree = ast.parse(open(path).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


@pytest.mark.skipif(_reference_functions("cccat") is None, reason="reference tree not mounted")
def test_cccat_and_clip_index_pinned_to_reference():
    ref_cccat, ref_index = _reference_functions("cccat", "get_original_index")
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 6, 3, 8, 8, generator=g)
    out = pipeline.cccat(a)
    assert out.shape == (2, 16, 3, 8, 8)
    assert torch.equal(out, ref_cccat(a))
    for world in (1, 2, 4, 8):
        for rank in range(world):
            for k, idx in enumerate(sharding.shard_indices(37, rank, world)):
                assert idx == ref_index(rank, k, interval=world) == sharding.original_index(rank, k, world)


def test_stage3_inputs_round_trip(tmp_path):
    clips = 5
    g = torch.Generator().manual_seed(1)
    key = torch.rand(clips, 3, 16, 16, generator=g)
    blurry = torch.rand(clips, 6, 3, 224, 224, generator=g)
    caps = [f"caption {i}" for i in range(clips)]
    torch.save(key, tmp_path / "video_subj02_all_recons.pt")
    torch.save(blurry.reshape(clips * 6, 3, 224, 224).half(), tmp_path / "recon_videos.pt")      # any leading shape / dtype: reshaped + .float() as :181
    torch.save(caps, tmp_path / "pred_test_caption.pt")
    inp = pipeline.Stage3Inputs.load(str(tmp_path), subj=2, clips=clips)
    assert len(inp) == clips and inp.blurry.shape == (clips, 6, 3, 224, 224) and inp.blurry.dtype == torch.float32
    seen = []
    for rank in range(2):
        for idx, k, b, c in inp.shard(rank, 2):
            assert torch.equal(k, key[idx]) and c == caps[idx] and b.shape == (6, 3, 224, 224)
            seen.append(idx)
    assert sorted(seen) == list(range(clips))
    with pytest.raises(ValueError):
        pipeline.Stage3Inputs.load(str(tmp_path), subj=2, clips=clips + 1)


def test_controlnet_condition_layout():
    imgs = torch.arange(2 * 4 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 4, 2, 3, 3)
    cond, mask = pipeline.controlnet_condition(imgs, [0, 5], 8)
    assert cond.shape == (2, 4, 8, 3, 3) and mask.shape == (2, 1, 8, 3, 3)
    assert torch.equal(cond[:, :, 0], imgs[:, :, 0]) and torch.equal(cond[:, :, 5], imgs[:, :, 1])
    assert cond[:, :, [1, 2, 3, 4, 6, 7]].abs().sum() == 0
    assert mask[:, :, [0, 5]].min() == 1 and mask.sum() == 2 * 2 * 9


@pytest.mark.skipif(not unet_shim.available(), reason="reference tree not mounted")
def test_denoiser_drives_reference_unet_and_controlnet():
    """The loop's noise predictor on the unmodified reference models (small widths, CPU): ControlNet residuals reach the UNet (the key
    frame changes the prediction), the call signature matches pipeline_neuroclips.py:464-475, and two loop steps run end to end."""
    small = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32, attention_head_dim=4, norm_num_groups=32)
    mm = dict(unet_shim.UNET_KW["motion_module_kwargs"], num_attention_heads=4)
    cmm = dict(unet_shim.CONTROLNET_KW["motion_module_kwargs"], num_attention_heads=4)
    with torch.no_grad():
        unet = unet_shim.build_unet(0, motion_module_kwargs=mm, **small)
        cn = unet_shim.build_controlnet(0, motion_module_kwargs=cmm, block_out_channels=(32, 64, 64, 64), cross_attention_dim=32, attention_head_dim=4)
        g = torch.Generator().manual_seed(2)
        for prm in cn.parameters():                  # ControlNet output / condition convolutions are zero-initialised (zero_module): randomise them
            if float(prm.abs().sum()) == 0.0:
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.05)
        lat = torch.randn(1, 4, 4, 8, 8, generator=g)
        ctx = torch.randn(2, 7, 32, generator=g)
        key = torch.randn(1, 4, 1, 8, 8, generator=g)
        x2 = torch.cat([lat] * 2)
        plain = pipeline.NeuroclipsDenoiser(unet)(x2, 961, ctx)
        d1 = pipeline.NeuroclipsDenoiser(unet, cn, key, (0,), 1.0)
        with_cn = d1(x2, 961, ctx)
        other = pipeline.NeuroclipsDenoiser(unet, cn, key * 2.0, (0,), 1.0)(x2, 961, ctx)
        assert with_cn.shape == x2.shape
        assert (with_cn - plain).abs().max() > 0 and (other - with_cn).abs().max() > 0
        # equals the call sequence of the reference loop written out by hand
        cond, mask = pipeline.controlnet_condition(key, (0,), 4)
        down, mid = cn(x2, 961, encoder_hidden_states=ctx, controlnet_cond=cond, conditioning_mask=mask, conditioning_scale=1.0, guess_mode=False,
                       return_dict=False)
        ref = unet(x2, 961, encoder_hidden_states=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid).sample
        assert torch.equal(ref, with_cn)
        out = pipeline.enhance_clip(d1, lat, ctx, sampler.DDIMSchedule(), num_inference_steps=2, guidance_scale=8.5, low_strength=0.3, seed=0)
        assert out.shape == lat.shape and torch.isfinite(out).all()
