"""SURVEY 8(f) N4 -- the blurry-video decoder's temporal attention + blend (model_variants/video_decoder.py:237-248,394-406) and the batched VAE
decode (pipeline_neuroclips.py:242-255).  PARITY UNPINNED for the attention block: diffusers' Attention class is neither installed nor
vendored, so the checker is the restated oracle (oracle/decoder_oracle.py) plus properties the reference code fixes by itself (weight = 1 is
the identity; the op is per position).  The batched decode IS pinned: against the reference's own decode_latents compiled from its file."""
import ast
import os
from types import SimpleNamespace

import pytest
import torch
from torch import nn

from oracle import decoder_oracle as do
from oracle import ref_shim
from tests.helpers import TOL_BF16, TOL_FP32, round_bf16


class FakeAttention(nn.Module):
    """Stand-in with the attribute names of diffusers' Attention that the block (and our patch) touch."""

    def __init__(self, cfg: do.DecoderAttnConfig, params):
        super().__init__()
        C = cfg.channels
        self.heads, self.rescale_output_factor, self.residual_connection = cfg.heads, cfg.rescale_output_factor, True
        self.spatial_norm, self.norm_cross = None, None
        self.group_norm = nn.GroupNorm(cfg.groups, C, eps=cfg.eps, affine=True)
        self.to_q, self.to_k, self.to_v = nn.Linear(C, C), nn.Linear(C, C), nn.Linear(C, C)
        self.to_out = nn.ModuleList([nn.Linear(C, C), nn.Dropout(0.0)])
        self.load_state_dict(params, strict=True)


def test_oracle_properties():
    cfg = do.DecoderAttnConfig(64)
    p = do.make_params(cfg, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2 * 6, 64, 5, 4, generator=g)
    assert torch.equal(do.temporal_blend_reference_order(p, x, 1.0, 6, cfg), x)            # weight = 1 (its init value, :217): identity
    y = do.temporal_blend_reference_order(p, x, 0.25, 6, cfg)
    assert y.shape == x.shape
    # per position: permuting the positions of the input permutes the output the same way
    perm = torch.randperm(20, generator=g)
    xp = x.reshape(12, 64, 20)[:, :, perm].reshape(12, 64, 5, 4)
    yp = do.temporal_blend_reference_order(p, xp, 0.25, 6, cfg)
    assert torch.allclose(yp.reshape(12, 64, 20), y.reshape(12, 64, 20)[:, :, perm], atol=1e-6)


def _reference_decode_latents():
    root = ref_shim.reference_root()
    path = os.path.join(root or "", "animatediff", "pipelines", "pipeline_neuroclips.py")
    if root is None or not os.path.isfile(path):
        return None
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and any(isinstance(m, ast.FunctionDef) and m.name == "decode_latents" for m in n.body))
    fn = next(m for m in cls.body if isinstance(m, ast.FunctionDef) and m.name == "decode_latents")
    from einops import rearrange
    ns = {"torch": torch, "rearrange": rearrange}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["decode_latents"]


@pytest.mark.skipif(_reference_decode_latents() is None, reason="reference tree not mounted")
def test_batched_decode_pinned_to_reference_decode_latents():
    from neurons_b200 import video_decoder as vd
    ref = _reference_decode_latents()
    g = torch.Generator().manual_seed(2)
    w = torch.randn(3, 4, generator=g)

    def toy_decode(z):                       # per-frame "VAE": a 1x1 convolution + 2x nearest up-sampling
        return torch.einsum("oc,nchw->nohw", w, z).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    fake_self = SimpleNamespace(vae=SimpleNamespace(decode=lambda z: SimpleNamespace(sample=toy_decode(z))))
    lat = torch.randn(2, 4, 16, 4, 4, generator=g)
    want = torch.from_numpy(ref(fake_self, lat))
    for chunk in (None, 5, 1):
        got = vd.decode_latents_batched(toy_decode, lat, chunk)
        assert got.shape == want.shape and torch.allclose(got, want, atol=1e-6)


def test_patch_has_no_cpu_path():
    from neurons_b200 import video_decoder as vd
    cfg = do.DecoderAttnConfig(32)
    blk = nn.Module()
    blk.attentions = nn.ModuleList([nn.Identity()])
    blk.temp_attentions = nn.ModuleList([FakeAttention(cfg, do.make_params(cfg, 0))])
    blk.resnets = nn.ModuleList([nn.Identity()])
    blk.weights = nn.ParameterList([nn.Parameter(torch.ones(1))])
    blk.upsamplers = None
    holder = nn.Module()
    holder.blk = blk
    assert vd.patch_video_decoder(holder) == 1
    with pytest.raises(RuntimeError):
        vd.temporal_attention_blend(torch.zeros(6, 32, 2, 2), blk.temp_attentions[0], blk.weights[0], 6)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("C,heads,b,side,weight", [(32, 1, 1, 7, 0.3), (64, 1, 2, 14, 0.0), (128, 1, 1, 28, 0.6), (128, 4, 1, 5, 0.3), (64, 1, 1, 3, 1.0)])
def test_decoder_temporal_attention_vs_oracle(dtype, C, heads, b, side, weight):
    from neurons_b200 import video_decoder as vd
    cfg = do.DecoderAttnConfig(C, heads)
    params = do.make_params(cfg, C + heads)
    g = torch.Generator().manual_seed(side)
    x = torch.randn(b * 6, C, side, side, generator=g)
    if dtype == torch.bfloat16:
        params, x = {k: round_bf16(v) for k, v in params.items()}, round_bf16(x)
    attn = FakeAttention(cfg, params).cuda().to(dtype).eval()
    wt = nn.Parameter(torch.tensor([weight]))
    with torch.no_grad():
        y = vd.temporal_attention_blend(x.cuda().to(dtype), attn, wt.cuda(), 6)
        ref = do.temporal_blend_reference_order(params, x, weight, 6, cfg)
    assert y.shape == x.shape and y.dtype == dtype
    err = (y.float().cpu() - ref).abs().max().item()
    assert err <= (TOL_FP32 if dtype == torch.float32 else TOL_BF16), err
    if weight == 1.0:
        assert torch.equal(y.float().cpu(), x)


@pytest.mark.gpu
def test_patched_blocks_match_reference_loop():
    """The patched block forwards against the reference loop written out with the oracle (resnets / spatial attentions are the block's own)."""
    from neurons_b200 import video_decoder as vd
    cfg = do.DecoderAttnConfig(64)
    params = [do.make_params(cfg, 10 + i) for i in range(2)]
    for mid in (False, True):
        blk = nn.Module()
        blk.attentions = nn.ModuleList([nn.Identity(), nn.Identity()])
        blk.temp_attentions = nn.ModuleList([FakeAttention(cfg, p) for p in params])
        blk.weights = nn.ParameterList([nn.Parameter(torch.tensor([0.4])), nn.Parameter(torch.tensor([0.7]))])

        class Res(nn.Module):
            def __init__(self, s):
                super().__init__()
                self.s = s

            def forward(self, h, temb=None):
                return h * self.s
        blk.resnets = nn.ModuleList([Res(1.1), Res(0.9), Res(1.05)] if mid else [Res(1.1), Res(0.9)])
        if not mid:
            blk.upsamplers = None
        for i in range(2):                    # the spatial attention modules take (h, temb=..., scale=...)
            blk.attentions[i].forward = (lambda h, temb=None, scale=1.0: h)
        holder = nn.Module()
        holder.blk = blk
        assert vd.patch_video_decoder(holder) == 1
        g = torch.Generator().manual_seed(5)
        x = torch.randn(6, 64, 6, 6, generator=g)
        with torch.no_grad():
            y = blk.cuda()(x.cuda(), None, time=6).cpu()
            h = x * 1.1
            if mid:
                for i in range(2):
                    h = do.temporal_blend_reference_order(params[i], h, float(blk.weights[i]), 6, cfg)
                    h = h * (0.9 if i == 0 else 1.05)
            else:
                h = do.temporal_blend_reference_order(params[0], h, 0.4, 6, cfg)
                h = do.temporal_blend_reference_order(params[1], h * 0.9, 0.7, 6, cfg)
        assert (y - h).abs().max().item() <= 5 * TOL_FP32
