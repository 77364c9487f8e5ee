"""SURVEY 8(f) row N1 as written: GroupNorm statistics carried between neighbours instead of recomputed.

  * the motion module's last kernel emits the per-(b, f, group) sums of its output y (proj_out's epilogue / the fused module's y store),
    so the next ResnetBlock3D.norm1 (resnet.py:182-198; call order unet_blocks.py:407-411) needs no statistics pass;
  * the module accepts the statistics of its own input from the producer of x.
Checker: torch in fp64 on the same tensors (plumbing only)."""
import pytest
import torch

import neurons_b200 as nb
from neurons_b200 import ops
from oracle import motion_oracle as mo
from tests import helpers

GPU = torch.cuda.is_available()
DEV = "cuda"


def _sums_ref(t):      # [b, c, f, h, w] -> float64 [B*F*32, 2]
    B, C, F, H, W = t.shape
    g = t.double().permute(0, 2, 1, 3, 4).reshape(B * F, 32, -1)
    return torch.stack([g.sum(-1), (g * g).sum(-1)], -1).reshape(B * F * 32, 2)


def test_carried_sums_are_dropped_when_the_tensor_changes():
    t = torch.zeros(1, 32, 2, 2, 2)
    s = torch.zeros(2 * 32, 2, dtype=torch.float64)
    assert nb.carried_sums(t) is None
    nb.attach_sums(t, s)
    assert nb.carried_sums(t) is s
    t.add_(1.0)                                   # in-place update bumps the version counter
    assert nb.carried_sums(t) is None
    nb.attach_sums(t, s)
    assert nb.carried_sums(t[:, :16]) is None      # a view is another tensor object


@pytest.mark.gpu
@pytest.mark.skipif(not GPU, reason="needs a CUDA device")
@pytest.mark.parametrize("C,F,H,W,B,dtype", [
    (320, 8, 16, 16, 2, torch.bfloat16),      # one-kernel module: statistics pass over y (default) / sums from its y store (option)
    (320, 16, 8, 8, 1, torch.bfloat16),
    (640, 8, 8, 8, 2, torch.bfloat16),        # multi-kernel path: sums from proj_out's epilogue
    (1280, 8, 8, 8, 1, torch.bfloat16),
    (320, 8, 3, 5, 1, torch.bfloat16),        # ragged latent: statistics pass over y
    (320, 8, 8, 8, 1, torch.float32),         # fp32 modes: statistics pass over y
])
@pytest.mark.parametrize("fused_emit", [0, 1])
def test_forward_stats_emits_sums_of_y_and_accepts_sums_of_x(C, F, H, W, B, dtype, fused_emit):
    from neurons_b200 import lib as nlib
    if fused_emit and not (C == 320 and dtype == torch.bfloat16 and (H * W) % (128 // F) == 0):
        pytest.skip("option only affects calls that run on the one-kernel path")
    with nlib.options({nlib.OPT_FUSED_Y_STATS: fused_emit}):
        _check_forward_stats(C, F, H, W, B, dtype)


def _check_forward_stats(C, F, H, W, B, dtype):
    cfg = mo.MotionConfig(C)
    params = mo.make_params(cfg, 5)
    x = mo.make_input((B, C, F, H, W), 6, layout="bfchw")
    m = helpers.mirror_module(cfg, params, DEV, dtype)
    xd = x.to(DEV, dtype)
    with torch.no_grad():
        y0 = m(xd, None, None)
        eng = m.__dict__["_nmm_engine"]
        y_sums = torch.full((B * F * 32, 2), float("nan"), dtype=torch.float64, device=DEV)
        y1 = ops.forward_packed(xd, eng.packed, eng.cfg, y_sums=y_sums)
        assert torch.equal(y0, y1)                                   # emitting the statistics does not change y
        ref = _sums_ref(y1.cpu())
        assert (y_sums.cpu() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
        # statistics of x handed in (here: computed by the library's own stand-alone kernel, then by torch in fp64)
        for x_sums in (ops.groupnorm_sums(xd), _sums_ref(xd.cpu()).to(DEV)):
            y2 = ops.forward_packed(xd, eng.packed, eng.cfg, x_sums=x_sums.contiguous())
            ulp = 2 ** -7 if dtype == torch.bfloat16 else 1e-5
            assert (y2.double() - y0.double()).abs().max().item() <= ulp * max(1.0, y0.double().abs().max().item())


@pytest.mark.gpu
@pytest.mark.skipif(not GPU, reason="needs a CUDA device")
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_inflated_groupnorm_with_carried_sums(dtype):
    B, C, F, H, W = 2, 640, 8, 8, 8
    x = mo.make_input((B, C, F, H, W), 9, layout="bfchw").to(DEV, dtype)
    g = torch.Generator().manual_seed(1)
    w, b = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV), (0.2 * torch.randn(C, generator=g)).to(DEV)
    y0 = ops.inflated_groupnorm(x, w, b, 1e-5, silu=True)
    y1 = ops.inflated_groupnorm(x, w, b, 1e-5, silu=True, sums=ops.groupnorm_sums(x))
    assert (y0.double() - y1.double()).abs().max().item() <= (2 ** -7 if dtype == torch.bfloat16 else 1e-5) * max(1.0, y0.double().abs().max().item())


@pytest.mark.gpu
@pytest.mark.skipif(not GPU, reason="needs a CUDA device")
def test_patched_stack_carries_the_statistics_from_module_to_norm():
    """motion module -> InflatedGroupNorm, both patched: with carry_stats the norm runs ONE kernel fewer (no statistics pass) and
    produces the same tensor."""
    C, B, F, H, W = 640, 2, 8, 8, 8

    class Stack(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mm = nb.get_motion_module(C, "Vanilla", dict(num_attention_heads=8, num_transformer_block=1,
                                                               attention_block_types=("Temporal_Self", "Temporal_Self"),
                                                               temporal_position_encoding=True, temporal_position_encoding_max_len=24,
                                                               temporal_attention_dim_div=1, zero_initialize=False))
            self.norm1 = nb.InflatedGroupNorm(32, C, eps=1e-5)

        def forward(self, x):
            return self.norm1(self.mm(x, None, None))

    torch.manual_seed(0)
    with torch.device(DEV):
        net = Stack().to(torch.bfloat16).eval()
    x = mo.make_input((B, C, F, H, W), 2, layout="bfchw").to(DEV, torch.bfloat16)
    from neurons_b200 import lib as nlib
    with torch.no_grad():
        assert nb.patch(net) == 1
        net(x)
        nlib.profile_begin(); y0 = net(x); plain = nlib.profile_end()
        assert nb.patch(net, carry_stats=True) == 1
        net(x)
        nlib.profile_begin(); y1 = net(x); carried = nlib.profile_end()
    # statistics passes (gn_stats launches): the module's own + the norm's -> the module's own only
    assert plain["gn_stats"]["launches"] == 2 and carried["gn_stats"]["launches"] == 1
    assert (y0.double() - y1.double()).abs().max().item() <= 2 ** -7 * max(1.0, y0.double().abs().max().item())
