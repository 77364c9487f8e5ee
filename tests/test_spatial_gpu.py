"""GPU parity of the spatial transformer through the C ABI: the flash-style attention kernel against an fp64 softmax
attention, the whole module against the reference-generated fixtures (tests/golden/sp_*.pt) in fp32 (bar 1e-4) and bf16 (bar 2e-2), and
-- at the UNet's full 64 x 64 size -- per-image independence: every (b, f) image of the big call equals the oracle on that image alone."""
import pytest
import torch

from oracle import spatial_oracle as so
from tests.helpers import TOL_BF16, TOL_FP32, round_bf16
from tests.test_spatial_oracle import load_sp_golden, sp_golden_names

pytestmark = pytest.mark.gpu


def _attention_fp64(q, k, v, heads, kv_div):
    """q [I, Lq, C], k / v [I / kv_div, Lkv, C] -> [I, Lq, C] in float64 (motion_module_new.py:258-287, heads folded)."""
    I, Lq, C = q.shape
    dh = C // heads
    kk = k.repeat_interleave(kv_div, dim=0).double().view(I, -1, heads, dh).permute(0, 2, 1, 3)
    vv = v.repeat_interleave(kv_div, dim=0).double().view(I, -1, heads, dh).permute(0, 2, 1, 3)
    qq = q.double().view(I, Lq, heads, dh).permute(0, 2, 1, 3)
    p = torch.softmax(qq @ kk.transpose(-1, -2) * dh ** -0.5, dim=-1)
    return (p @ vv).permute(0, 2, 1, 3).reshape(I, Lq, C)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("dh", [40, 80, 160])
@pytest.mark.parametrize("Lq,Lkv,images,kv_div", [(64, 64, 3, 1), (192, 192, 2, 1), (1024, 1024, 1, 1), (200, 77, 4, 2), (20, 20, 2, 1), (300, 1, 1, 1),
                                                  (300, 300, 3, 1), (513, 513, 2, 1), (256, 256, 5, 1)])      # ragged / odd tile counts on the tcgen05 kernel
def test_spatial_attention_vs_fp64(dtype, dh, Lq, Lkv, images, kv_div):
    import neurons_b200 as nb
    if dtype == torch.float32 and Lq * Lkv > 200 * 200:
        pytest.skip("fp32 checker kernel: small shapes only")
    heads = 8 if dh < 160 else 4
    C = heads * dh
    g = torch.Generator().manual_seed(Lq * 7 + Lkv + dh)
    if kv_div == 1 and Lq == Lkv:                    # self-attention: q | k | v are column slices of one [I, L, 3C] projection output
        qkv = torch.randn(images, Lq, 3 * C, generator=g).to(dtype).cuda()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:                                            # cross-attention: k | v slices of a [I / kv_div, Lkv, 2C] buffer
        q = torch.randn(images, Lq, C, generator=g).to(dtype).cuda()
        kvb = torch.randn(images // kv_div, Lkv, 2 * C, generator=g).to(dtype).cuda()
        k, v = kvb[..., :C], kvb[..., C:]
    o = nb.spatial_attention(q, k, v, heads, kv_div)
    torch.cuda.synchronize()
    ref = _attention_fp64(q.cpu(), k.cpu(), v.cpu(), heads, kv_div)
    err = (o.double().cpu() - ref).abs().max().item()
    assert err <= (1.5e-2 if dtype == torch.bfloat16 else 2e-5), err


def _mirror(cfg, params, dtype):
    import neurons_b200 as nb
    m = nb.Transformer3DModel(num_attention_heads=cfg.heads, attention_head_dim=cfg.head_dim, in_channels=cfg.channels, num_layers=cfg.layers,
                              cross_attention_dim=cfg.ctx_dim, use_linear_projection=not cfg.conv_proj, unet_use_cross_frame_attention=False,
                              unet_use_temporal_attention=False)
    m.load_state_dict(params, strict=True)
    return m.eval().cuda().to(dtype)


@pytest.mark.parametrize("mode", ["x3", "fma"])          # Linears on the tensor cores (3 x bf16 per product) / on the FMA pipe (the checker)
@pytest.mark.parametrize("name", sp_golden_names())
def test_spatial_module_golden_fp32(name, mode):
    fx, cfg, params, x, ctx = load_sp_golden(name)
    with torch.no_grad():
        m = _mirror(cfg, params, torch.float32)
        if mode == "fma":
            m.__dict__["_nmm_fp32_fma"] = True
        y = m(x.cuda(), encoder_hidden_states=ctx.cuda()).sample
    assert y.shape == fx["out_ref_fp32"].shape and y.stride() == tuple(fx["out_ref_fp32"].permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4).stride())
    err = (y.cpu() - fx["out_ref_fp32"]).abs().max().item()
    assert err <= TOL_FP32, err


@pytest.mark.parametrize("name", sp_golden_names())
def test_spatial_module_golden_bf16(name):
    fx, cfg, params, x, ctx = load_sp_golden(name)
    with torch.no_grad():
        y = _mirror(cfg, params, torch.bfloat16)(x.cuda().bfloat16(), encoder_hidden_states=ctx.cuda().bfloat16()).sample
    assert y.dtype == torch.bfloat16
    err = (y.float().cpu() - fx["out_ref_bf16in"]).abs().max().item()
    assert err <= TOL_BF16, err


@pytest.mark.parametrize("C,side,frames", [(320, 64, 8), (640, 32, 8), (1280, 16, 16)])
def test_spatial_full_size_image_independence(C, side, frames):
    """UNet-level sizes (config 2: CFG batch 2, 8 frames, 64 x 64 latent at C = 320): the oracle cannot run the whole tensor in seconds,
    but every (b, f) image of the module is independent of the others (GroupNorm, both attentions and every Linear are per image), so
    one image of the full-size result must equal the oracle on that image alone."""
    cfg = so.SpatialConfig(C, 8, 1, 768, True)
    params = so.make_params(cfg, 31)
    B = 2
    x, ctx = so.make_inputs(cfg, B, frames, side, side, 77, 32)
    pb = {k: round_bf16(v) for k, v in params.items()}
    with torch.no_grad():
        y = _mirror(cfg, params, torch.bfloat16)(x.cuda().bfloat16(), encoder_hidden_states=ctx.cuda().bfloat16()).sample
        torch.cuda.synchronize()
        assert torch.isfinite(y.float()).all()
        for b, f in ((0, 0), (1, frames - 1)):
            ref = so.forward_reference_order(pb, round_bf16(x[b:b + 1, :, f:f + 1]), round_bf16(ctx[b:b + 1]), cfg)
            err = (y[b:b + 1, :, f:f + 1].float().cpu() - ref).abs().max().item()
            assert err <= TOL_BF16, (b, f, err)


@pytest.mark.parametrize("C,side,frames", [(320, 64, 2), (640, 32, 4)])
def test_spatial_full_size_fp32_tensor_core_mode(C, side, frames):
    """fp32 activations (the reference as shipped) at the UNet's latent sizes: Linears and attention on the tensor cores as 3 bf16 MMAs per
    product (NMM_F32X3; the attention takes q, k, v and the softmax weights as hi | lo bf16 splits) -- one image against the fp32 oracle on that
    image alone, bar 1e-4."""
    cfg = so.SpatialConfig(C, 8, 1, 768, True)
    params = so.make_params(cfg, 41)
    x, ctx = so.make_inputs(cfg, 1, frames, side, side, 77, 42)
    with torch.no_grad():
        m = _mirror(cfg, params, torch.float32)
        y = m(x.cuda(), encoder_hidden_states=ctx.cuda()).sample
        torch.cuda.synchronize()
        f = frames - 1
        ref = so.forward_reference_order(params, x[:, :, f:f + 1], ctx, cfg)
        err = (y[:, :, f:f + 1].cpu() - ref).abs().max().item()
        assert err <= TOL_FP32, err


def test_spatial_then_motion_chain():
    """The call order of every CrossAttn block (unet_blocks.py:409-411): spatial transformer -> motion module, the second consuming the
    first's [B,F,C,H,W]-storage view directly."""
    import neurons_b200 as nb
    from oracle import motion_oracle as mo
    from tests.helpers import mirror_module
    cfg = so.SpatialConfig(320, 8, 1, 768, True)
    params = so.make_params(cfg, 5)
    x, ctx = so.make_inputs(cfg, 1, 8, 8, 8, 77, 6)
    mcfg = mo.MotionConfig(320, 8, 1, 2, True, 24)
    mparams = mo.make_params(mcfg, 7)
    with torch.no_grad():
        sp = _mirror(cfg, params, torch.float32)
        mm = mirror_module(mcfg, mparams, device="cuda")
        n0 = nb.launch_count()
        y = mm(sp(x.cuda(), encoder_hidden_states=ctx.cuda()).sample, None, None)
        torch.cuda.synchronize()
        assert nb.launch_count() > n0
        ref = mo.forward_reference_order(mparams, so.forward_reference_order(params, x, ctx, cfg), mcfg)
    assert (y.cpu() - ref).abs().max().item() <= 2 * TOL_FP32


@pytest.mark.parametrize("C,side", [(640, 16), (320, 16)])
def test_spatial_emits_groupnorm_sums_for_the_motion_module(C, side):
    """SURVEY 8(f) N1 with its real producer: patch_spatial(carry_stats=True) makes proj_out's epilogue emit the GroupNorm sums of the
    spatial transformer's output; the motion module behind it (patch(carry_stats=True)) takes them instead of running its own statistics
    pass.  Same result (the sums are those of y as stored), one gn_stats launch fewer per block."""
    import torch.nn as nn
    import neurons_b200 as nb
    from neurons_b200 import lib as nlib
    from oracle import motion_oracle as mo
    from tests.helpers import mirror_module
    cfg = so.SpatialConfig(C, 8, 1, 768, True)
    params = {k: round_bf16(v) for k, v in so.make_params(cfg, 3).items()}
    x, ctx = so.make_inputs(cfg, 2, 8, side, side, 77, 4)
    mcfg = mo.MotionConfig(C, 8, 1, 2, True, 24)
    mparams = mo.make_params(mcfg, 5)
    with torch.no_grad():
        holder = nn.Module()
        holder.sp = _mirror(cfg, params, torch.bfloat16)
        holder.mm = mirror_module(mcfg, mparams, device="cuda", dtype=torch.bfloat16)
        xb, cb = x.cuda().bfloat16(), ctx.cuda().bfloat16()

        def run():
            nlib.profile_begin()
            y = holder.mm(holder.sp(xb, encoder_hidden_states=cb).sample, None, None)
            prof = nlib.profile_end()
            return y.float().cpu(), prof["gn_stats"]["launches"]
        assert nb.patch(holder, carry_stats=True) == 1        # the motion module emits the sums of ITS output in both runs (a statistics pass
        y0, n0 = run()                                        # on the one-kernel C = 320 path, proj_out's epilogue elsewhere)
        assert nb.patch_spatial(holder, carry_stats=True) == 1
        y1, n1 = run()
        sums = nb.carried_sums(holder.sp(xb, encoder_hidden_states=cb).sample)
        ref = nb.ops.groupnorm_sums(holder.sp(xb, encoder_hidden_states=cb).sample)
    assert n1 == n0 - 1, (n0, n1)           # the motion module's statistics pass over its input is gone
    assert sums is not None and torch.allclose(sums, ref, rtol=1e-5, atol=1e-3)
    assert (y1 - y0).abs().max().item() <= 2.0 ** -6


def test_spatial_forward_graph_capture_and_errors():
    """The spatial call allocates nothing, never synchronises and encodes its tensor maps on the host: it can be captured in a CUDA graph
    and replayed on new data; a packed buffer of the wrong size, a missing text tensor and a CPU tensor raise instead of falling back."""
    import neurons_b200 as nb
    from neurons_b200 import spatial_transformer as st
    cfg = so.SpatialConfig(320, 8, 1, 768, True)
    params = {k: round_bf16(v) for k, v in so.make_params(cfg, 8).items()}
    m = _mirror(cfg, params, torch.bfloat16)
    x, ctx = so.make_inputs(cfg, 1, 2, 16, 16, 77, 9)
    xs, cs = x.cuda().bfloat16(), ctx.cuda().bfloat16()
    with torch.no_grad():
        y_eager = m(xs, encoder_hidden_states=cs).sample.clone()           # also packs the parameters (required before capture)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(xs, encoder_hidden_states=cs)                                 # warm-up on the capture stream
            with torch.cuda.graph(g, stream=s):
                y_graph = m(xs, encoder_hidden_states=cs).sample
        torch.cuda.current_stream().wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y_graph, y_eager)
        x2, _ = so.make_inputs(cfg, 1, 2, 16, 16, 77, 10)
        xs.copy_(x2.cuda().bfloat16())                                       # new data in the captured input buffer
        g.replay()
        torch.cuda.synchronize()
        ref = so.forward_reference_order(params, round_bf16(x2), round_bf16(ctx), cfg)
        assert (y_graph.float().cpu() - ref).abs().max().item() <= TOL_BF16
        eng = m.__dict__["_nmm_engine"]
        with pytest.raises(nb.NmmError):
            st.spatial_forward_packed(xs, cs, eng.packed[:-256], eng.cfg)
        with pytest.raises(ValueError):
            m(xs, encoder_hidden_states=None)
        with pytest.raises(RuntimeError):
            m(xs.cpu(), encoder_hidden_states=cs.cpu())
        with pytest.raises(TypeError):
            m(xs, encoder_hidden_states=cs.float())
