"""GPU (-m gpu): the CUDA path through the C ABI against the CPU oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): fp32 mode max-abs <= 1e-4 vs the fp32 reference; bf16 mode max-abs <= 2e-2 vs the
reference evaluated on the same bf16-rounded inputs and weights.  Stage-level tests use tighter, stage-appropriate
tolerances written at each assert.
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


@pytest.fixture(scope="module", autouse=True)
def _require_b200(built_library):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    with torch.cuda.device(0):
        rc = built_library.nmm_device_check()
    assert rc == 0, built_library.nmm_last_error()


def _cfg(c: mo.MotionConfig) -> nb.ModuleConfig:
    return nb.ModuleConfig(c.channels, c.heads, c.layers, c.attn_blocks, c.pos_enc, c.max_len)


def _maxabs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


SHAPES = [  # (C, F, H, W, B, layout)
    (320, 8, 16, 16, 1, "bcfhw"),
    (320, 16, 8, 8, 2, "bfchw"),
    (640, 8, 8, 8, 1, "bfchw"),
    (64, 5, 3, 5, 2, "bcfhw"),        # ragged: odd P, C not a multiple of 64
    (1280, 16, 2, 2, 1, "bfchw"),
    (32, 1, 1, 2, 1, "bcfhw"),
]


@pytest.mark.parametrize("C,F,H,W,B,layout", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_groupnorm_stats_and_tokens(C, F, H, W, B, layout, dtype):
    cfg = mo.MotionConfig(C)
    params = mo.make_params(cfg, 1)
    x = mo.make_input((B, C, F, H, W), 2, layout=layout)
    if dtype == torch.bfloat16:
        x = helpers.round_bf16(x)
    st = mo.forward_token_order(params, x, cfg, torch.float64)
    xd = x.to(DEV, dtype)
    assert xd.stride() == x.stride()
    mean, rstd = ops.groupnorm_stats(_cfg(cfg), xd)
    assert _maxabs(mean, st.gn_mean) <= 2e-6
    assert (rstd.double().cpu() / st.gn_rstd - 1).abs().max().item() <= 5e-6
    gw, gb = params["temporal_transformer.norm.weight"].to(DEV), params["temporal_transformer.norm.bias"].to(DEV)
    tok = ops.groupnorm_tokens(_cfg(cfg), xd, gw, gb)
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7 * 1.01 * st.tokens.abs().max().item() / 2     # half-ulp of bf16 at the max
    assert _maxabs(tok, st.tokens) <= tol


@pytest.mark.parametrize("C,F,H,W,B,layout", [(320, 8, 8, 8, 1, "bcfhw"), (320, 8, 16, 16, 2, "bfchw"), (640, 16, 8, 8, 1, "bfchw"),
                                              (1280, 8, 8, 8, 2, "bfchw"), (96, 2, 8, 8, 1, "bcfhw")])
def test_groupnorm_linear_fused_matches_two_kernel_path(C, F, H, W, B, layout):
    """GroupNorm + re-layout + proj_in in one tcgen05 kernel (x tiles TMA-loaded as an M-major operand, normalised in shared
    memory) against gn_tokens + linear (bit-identical operands -> same accumulation) and against the fp64 oracle."""
    cfg = mo.MotionConfig(C)
    params = mo.make_params(cfg, 1)
    x = helpers.round_bf16(mo.make_input((B, C, F, H, W), 2, layout=layout))
    st = mo.forward_token_order(params, x, cfg, torch.float64)
    xd = x.to(DEV, torch.bfloat16)
    gw, gb = params["temporal_transformer.norm.weight"].to(DEV), params["temporal_transformer.norm.bias"].to(DEV)
    w = helpers.round_bf16(params["temporal_transformer.proj_in.weight"]).to(DEV, torch.bfloat16)
    b = params["temporal_transformer.proj_in.bias"].to(DEV)
    fused = ops.groupnorm_linear(_cfg(cfg), xd, gw, gb, w, b)
    tok = ops.groupnorm_tokens(_cfg(cfg), xd, gw, gb)
    h2 = torch.empty_like(fused)
    ops.linear(tok, w, b, nlib.EPI_STORE, h=h2, want_out=False)
    assert torch.equal(fused, h2)
    ref = st.tokens @ w.double().cpu().T + b.double().cpu()
    assert _maxabs(fused, ref) <= 2e-2


@pytest.mark.parametrize("C,F,H,W,B", [(320, 8, 8, 8, 1), (640, 16, 4, 4, 1), (1280, 8, 2, 2, 2), (96, 3, 3, 3, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("with_pe", [True, False])
def test_layernorm_pe(C, F, H, W, B, dtype, with_pe):
    cfg = mo.MotionConfig(C)
    g = torch.Generator().manual_seed(5)
    N = B * F * H * W
    h = torch.randn(N, C, generator=g) * 1.7 + 0.3
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    pe = mo.positional_encoding(cfg.max_len, C)
    ref = torch.nn.functional.layer_norm(h.double(), (C,), w.double(), b.double(), 1e-5)
    if with_pe:
        f_of_n = (torch.arange(N) // (H * W)) % F
        ref = ref + pe.double()[f_of_n]
    out = ops.layernorm_pe(_cfg(cfg), (B, F, H, W), h.to(DEV), w.to(DEV), b.to(DEV), pe.to(DEV) if with_pe else None, dtype)
    tol = 1e-5 if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item()
    assert _maxabs(out, ref) <= tol


@pytest.mark.parametrize("C,F,H,W,B", [(320, 8, 8, 8, 1), (320, 16, 4, 5, 2), (640, 24, 2, 2, 1), (1280, 16, 2, 2, 1),
                                       (32, 1, 1, 2, 1), (64, 32, 2, 2, 1), (1280, 8, 4, 4, 1),
                                       # ragged position counts (last tile of the specialised kernel holds fewer positions), every d_h x F variant
                                       (640, 8, 3, 5, 2), (1280, 8, 3, 3, 1), (640, 16, 3, 3, 2), (320, 8, 3, 5, 1), (320, 8, 33, 33, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_temporal_attention(C, F, H, W, B, dtype):
    cfg = mo.MotionConfig(C, max_len=32)
    nh, dh, P = cfg.heads, cfg.head_dim, H * W
    N = B * F * P
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(N, 3 * C, generator=g)
    if dtype == torch.bfloat16:
        qkv = helpers.round_bf16(qkv)
    z = qkv.double().reshape(B, F, P, 3, nh, dh)
    s = torch.einsum("bfphd,bgphd->bphfg", z[:, :, :, 0], z[:, :, :, 1]) * dh ** -0.5
    ref = torch.einsum("bphfg,bgphd->bfphd", s.softmax(-1), z[:, :, :, 2]).reshape(N, C)
    ctx = ops.temporal_attention(_cfg(cfg), (B, F, H, W), qkv.to(DEV, dtype))
    tol = 2e-5 if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item() + 1e-5
    assert _maxabs(ctx, ref) <= tol


GEMM_SHAPES = [  # (M, N, K)
    (256, 320, 320), (1000, 960, 320), (300, 2560, 320), (130, 320, 1280), (64, 64, 64), (2, 32, 32), (513, 96, 128),
    (4096, 640, 640), (128, 1280, 5120),
]


def _gemm_inputs(M, N, K, dtype, seed=3):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    if dtype == torch.bfloat16:
        A, W = helpers.round_bf16(A), helpers.round_bf16(W)
    return A, W, bias


def _gemm_tol(dtype, ref, K):
    # fp32 accumulate of exactly-representable products: only summation-order error; bf16 outputs add one rounding
    return 5e-5 if dtype == torch.float32 else 2e-4


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_linear_store_and_residual(M, N, K, dtype):
    A, W, bias = _gemm_inputs(M, N, K, dtype)
    ref = A.double() @ W.double().T + bias.double()
    Ad, Wd, bd = A.to(DEV, dtype), W.to(DEV, dtype), bias.to(DEV)
    # STORE into the fp32 residual stream (exact fp32 accumulate, no output rounding)
    h = torch.full((M, N), float("nan"), device=DEV)
    out = ops.linear(Ad, Wd, bd, nlib.EPI_STORE, h=h, want_out=True)
    assert _maxabs(h, ref) <= _gemm_tol(dtype, ref, K)
    out_tol = _gemm_tol(dtype, ref, K) if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item()
    assert _maxabs(out, ref) <= out_tol
    # no-bias STORE (QKV projection)
    out2 = ops.linear(Ad, Wd, None, nlib.EPI_STORE)
    assert _maxabs(out2, ref - bias.double()) <= out_tol
    # RESIDUAL: h = acc + bias + h, in place ...
    g = torch.Generator().manual_seed(11)
    h0 = torch.randn(M, N, generator=g)
    h = h0.to(DEV).clone()
    assert ops.linear(Ad, Wd, bd, nlib.EPI_RESIDUAL, h=h, want_out=False) is None
    assert _maxabs(h, ref + h0.double()) <= _gemm_tol(dtype, ref, K)
    # ... or, with `out`, the sum goes to out only and h is left untouched
    h = h0.to(DEV).clone()
    out3 = ops.linear(Ad, Wd, bd, nlib.EPI_RESIDUAL, h=h, want_out=True)
    assert torch.equal(h.cpu(), h0)
    assert _maxabs(out3, ref + h0.double()) <= (out_tol if dtype == torch.float32 else 2 ** -8 * 1.01 * (ref + h0.double()).abs().max().item())


@pytest.mark.parametrize("M,N,K", [(256, 2560, 320), (130, 512, 64), (2, 256, 32), (1024, 5120, 640)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_linear_geglu(M, N, K, dtype):
    # library contract: W rows pre-interleaved in groups of four (value 2q, value 2q+1, gate 2q, gate 2q+1); emulate the packing here
    A, W, bias = _gemm_inputs(M, N, K, dtype)
    half = N // 2
    u = A.double() @ W.double().T + bias.double()
    ref = u[:, :half] * torch.nn.functional.gelu(u[:, half:])       # value * gelu(gate), motion_module_new.py:516-518
    Wi = torch.empty_like(W); Wi[0::4] = W[0:half:2]; Wi[1::4] = W[1:half:2]; Wi[2::4] = W[half::2]; Wi[3::4] = W[half + 1::2]
    bi = torch.empty_like(bias); bi[0::4] = bias[0:half:2]; bi[1::4] = bias[1:half:2]; bi[2::4] = bias[half::2]; bi[3::4] = bias[half + 1::2]
    out = ops.linear(A.to(DEV, dtype), Wi.to(DEV, dtype), bi.to(DEV), nlib.EPI_GEGLU)
    assert out.shape == (M, half)
    tol = 5e-5 if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item() + 2e-4
    assert _maxabs(out, ref) <= tol


@pytest.mark.parametrize("C,F,H,W,B,layout", [(320, 8, 8, 8, 1, "bcfhw"), (64, 5, 3, 5, 2, "bfchw"), (640, 16, 4, 4, 1, "bfchw")])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_linear_output_epilogue(C, F, H, W, B, layout, dtype):
    cfg = mo.MotionConfig(C)
    N = B * F * H * W
    A, Wt, bias = _gemm_inputs(N, C, C, dtype)
    x = mo.make_input((B, C, F, H, W), 4, layout=layout)
    if dtype == torch.bfloat16:
        x = helpers.round_bf16(x)
    y_tok = A.double() @ Wt.double().T + bias.double()                                   # [N, C], n = (b f) p
    ref = y_tok.reshape(B, F, H * W, C).permute(0, 3, 1, 2).reshape(B, C, F, H, W) + x.double()
    y = ops.linear(A.to(DEV, dtype), Wt.to(DEV, dtype), bias.to(DEV), nlib.EPI_OUTPUT, cfg=_cfg(cfg), x=x.to(DEV, dtype))
    assert y.shape == x.shape and y.permute(0, 2, 1, 3, 4).is_contiguous()                # [B,F,C,H,W] storage like the reference
    tol = 5e-5 if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item()
    assert _maxabs(y, ref) <= tol


# ---- whole module ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["tc", "fma"])
@pytest.mark.parametrize("name", helpers.golden_names())
def test_module_golden_fp32(name, mode):
    """fp32 activations, both arithmetic modes: "tc" = every Linear as 3 bf16 tcgen05 MMAs per product (NMM_F32X3, the default when
    channels % 64 == 0; narrower modules run the FMA path either way), "fma" = the fp32 FMA-pipe GEMM (NMM_F32, the checker)."""
    fx, cfg, params, x = helpers.load_golden(name)
    m = helpers.mirror_module(cfg, params, DEV)
    if mode == "fma":
        m.__dict__["_nmm_fp32_fma"] = True
    n0 = nlib.launch_count()
    with torch.no_grad():
        y = m(x.to(DEV), None, None)
    assert y.shape == x.shape and y.dtype == torch.float32
    assert y.stride() == fx["out_ref_fp32"].permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4).stride()
    assert _maxabs(y, fx["out_ref_fp32"]) <= helpers.TOL_FP32
    assert nlib.launch_count() > n0


def test_fp32_tc_and_fma_modes_agree_and_differ_in_kernels():
    """The two fp32 modes are different kernels (tcgen05 vs FMA) computing the same function: results agree far inside the bar, and
    the profile shows which GEMM ran."""
    fx, cfg, params, x = helpers.load_golden("c320_f8_8x8_a2")
    ys, kernels = [], []
    for fma in (False, True):
        m = helpers.mirror_module(cfg, params, DEV)
        if fma:
            m.__dict__["_nmm_fp32_fma"] = True
        with torch.no_grad():
            m(x.to(DEV), None, None)                       # packs
            nlib.profile_begin()
            ys.append(m(x.to(DEV), None, None))
            prof = nlib.profile_end()
        kernels.append({k for k, v in prof.items() if v["launches"] > 0})
    assert "linear_bf16_tcgen05" in kernels[0] and "linear_fp32_fma" not in kernels[0]
    assert "linear_fp32_fma" in kernels[1] and "linear_bf16_tcgen05" not in kernels[1]
    assert _maxabs(ys[0], ys[1].double()) <= 5e-5


@pytest.mark.parametrize("name", helpers.golden_names())
def test_module_golden_bf16(name):
    fx, cfg, params, x = helpers.load_golden(name)
    m = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)        # weights rounded to bf16 == fixture's rounding
    with torch.no_grad():
        y = m(x.to(DEV, torch.bfloat16), None, None)
    assert y.dtype == torch.bfloat16
    assert _maxabs(y, fx["out_ref_bf16in"]) <= helpers.TOL_BF16


def test_module_bf16_from_fp32_weights_matches_bf16_weights():
    # packing converts fp32 parameters to bf16 with the same round-to-nearest as module.bfloat16()
    fx, cfg, params, x = helpers.load_golden("c320_f8_8x8_a2")
    xb = x.to(DEV, torch.bfloat16)
    with torch.no_grad():
        y1 = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)(xb, None, None)
        packed = ops.pack_params(_cfg(cfg), {k: v.to(DEV) for k, v in params.items()}, torch.bfloat16, torch.device(DEV))
        y2 = ops.forward_packed(xb, packed, _cfg(cfg))
    # pe table: module.bfloat16() rounds the buffer, the fp32 source does not -> results may differ by
    # one bf16 ulp of the output where the pre-rounding values straddle a rounding boundary
    assert _maxabs(y1, y2) <= 2 ** -7 * 1.01 * y1.float().abs().max().item()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, helpers.TOL_FP32), (torch.bfloat16, helpers.TOL_BF16)], ids=["fp32", "bf16"])
def test_module_config1_full_size_vs_oracle(dtype, tol):
    """BASELINE config 1: 320 ch, 8 frames, 64x64 latent, batch 1 -- CUDA vs the oracle run live on the host CPU."""
    cfg = mo.MotionConfig(320)
    params = mo.make_params(cfg, 0)
    x = mo.make_input((1, 320, 8, 64, 64), 0)
    if dtype == torch.bfloat16:
        params = {k: helpers.round_bf16(v) for k, v in params.items()}
        x = helpers.round_bf16(x)
    with torch.no_grad():
        ref = mo.forward_reference_order(params, x, cfg)
        y = helpers.mirror_module(cfg, params, DEV, dtype)(x.to(DEV, dtype), None, None)
    assert _maxabs(y, ref) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_module_properties_full_size(dtype):
    """Size-independent properties at a NEURONS-sized call (B=2, 16 frames, 32x32 latent)."""
    cfg = mo.MotionConfig(320)
    params = mo.make_params(cfg, 4)
    x = mo.make_input((2, 320, 16, 32, 32), 6, layout="bfchw").to(DEV, dtype)
    m = helpers.mirror_module(cfg, params, DEV, dtype)
    with torch.no_grad():
        y1 = m(x, None, None)
        y2 = m(x, None, None)
        assert torch.equal(y1, y2)                                             # deterministic (no atomics anywhere)
        yc = m(x.contiguous(), None, None)                                     # layout of x does not change the result
        assert torch.equal(y1, yc)
        # frames are coupled only through attention; batch entries are independent
        y_b0 = m(x[:1], None, None)
        assert torch.equal(y_b0, y1[:1])
        # zero-initialised proj_out (the construction default, motion_module.py:74-75) makes the module the identity
        nb.zero_module(m.temporal_transformer.proj_out)
        nb.invalidate(m)
        assert torch.equal(m(x, None, None), x)


def test_layernorm_folding_matches_separate_layernorm():
    # ln_fold folds the LayerNorms into the QKV / GEGLU GEMMs (3 launches fewer per call); the default runs the
    # separate LayerNorm kernel.  Both variants must meet the bar and agree closely.  (Multi-kernel pipeline: fused module off.)
    fx, cfg, params, x = helpers.load_golden("c320_f16_8x8_a2_view")
    xb = x.to(DEV, torch.bfloat16)
    with torch.no_grad(), nlib.options({nlib.OPT_FUSED_MODULE: 0}):
        y_sep = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)(xb, None, None)
        m2 = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)
        m2.__dict__["_nmm_ln_fold"] = True
        m2(xb, None, None)
        n0 = nb.launch_count()
        y_fold = m2(xb, None, None)
        assert nb.launch_count() - n0 == 12
    assert _maxabs(y_fold, fx["out_ref_bf16in"]) <= helpers.TOL_BF16
    assert _maxabs(y_sep, fx["out_ref_bf16in"]) <= helpers.TOL_BF16
    assert _maxabs(y_fold, y_sep) <= 2 ** -7 * 1.01 * y_sep.float().abs().max().item()      # one output ulp


def test_custom_op_matches_module_call():
    # torch.ops.neurons_mm.forward (CUDA dispatch key) is the same computation the module's forward reaches directly
    fx, cfg, params, x = helpers.load_golden("c64_f16_4x4_a1_view")
    m = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)
    xb = x.to(DEV, torch.bfloat16)
    with torch.no_grad():
        y_mod = m(xb, None, None)
        packed = ops.pack_params(_cfg(cfg), {k: v for k, v in nb.motion_module._param_tensors(m).items()
                                             if not k.endswith("pos_encoder.pe")} |
                                 {k: v[0] for k, v in nb.motion_module._param_tensors(m).items() if k.endswith("pos_encoder.pe")},
                                 torch.bfloat16, torch.device(DEV))
        y_op = torch.ops.neurons_mm.forward(xb, packed, cfg.channels, cfg.heads, cfg.layers, cfg.attn_blocks, cfg.pos_enc, cfg.max_len)
    assert torch.equal(y_mod, y_op)
    assert _maxabs(y_op, fx["out_ref_bf16in"]) <= helpers.TOL_BF16


def test_lora_style_inplace_update_needs_invalidate():
    fx, cfg, params, x = helpers.load_golden("c64_f16_4x4_a1_view")
    m = helpers.mirror_module(cfg, params, DEV)
    xd = x.to(DEV)
    with torch.no_grad():
        y0 = m(xd, None, None)
        m.temporal_transformer.proj_out.weight.data += 0.05       # in place, like convert_lora_safetensor_to_diffusers.py:45
        assert torch.equal(m(xd, None, None), y0)                  # cached pack (documented behaviour)
        nb.invalidate(m)
        y1 = m(xd, None, None)
    assert not torch.equal(y1, y0)
    p2 = dict(params); p2["temporal_transformer.proj_out.weight"] = params["temporal_transformer.proj_out.weight"] + 0.05
    assert _maxabs(y1, mo.forward_reference_order(p2, x, cfg)) <= helpers.TOL_FP32


def test_inplace_updates_through_the_tensor_repack_automatically():
    """Updates that bump tensor._version -- load_state_dict / copy_ (load_weights, animatediff/utils/util.py:107-121), `p += d`,
    optimizer steps -- are seen by the packed-parameter cache without invalidate(); only `.data` writes need it (test above)."""
    fx, cfg, params, x = helpers.load_golden("c64_f16_4x4_a1_view")
    m = helpers.mirror_module(cfg, params, DEV)
    xd = x.to(DEV)
    with torch.no_grad():
        y0 = m(xd, None, None)
        p2 = {k: v.clone() for k, v in params.items()}
        p2["temporal_transformer.proj_out.weight"] += 0.05
        m.load_state_dict(p2, strict=False)                        # copy_ into the existing parameters: same data_ptr, new _version
        y1 = m(xd, None, None)
        assert _maxabs(y1, mo.forward_reference_order(p2, x, cfg)) <= helpers.TOL_FP32
        m.temporal_transformer.proj_out.bias += 0.25               # in place through the tensor itself
        p2["temporal_transformer.proj_out.bias"] += 0.25
        y2 = m(xd, None, None)
        assert _maxabs(y2, mo.forward_reference_order(p2, x, cfg)) <= helpers.TOL_FP32
    assert not torch.equal(y1, y0) and not torch.equal(y2, y1)


def test_packed_buffer_mismatch_is_refused():
    """nmm_forward checks the packed buffer's size against the call's layout (a buffer packed for another dtype / heads
    must not be read out of bounds) and the buffer starts with a header naming what it was packed for."""
    cfg = mo.MotionConfig(64)
    params = {k: v.to(DEV) for k, v in mo.make_params(cfg, 1).items()}
    packed32 = ops.pack_params(_cfg(cfg), params, torch.float32, torch.device(DEV))
    packed16 = ops.pack_params(_cfg(cfg), params, torch.bfloat16, torch.device(DEV))
    assert bytes(packed32[:64].cpu().numpy().tobytes()) == ops.packed_header(_cfg(cfg), torch.float32)
    assert bytes(packed16[:64].cpu().numpy().tobytes()) == ops.packed_header(_cfg(cfg), torch.bfloat16)
    x = torch.zeros(1, 64, 4, 2, 2, device=DEV)
    with pytest.raises(nlib.NmmError, match="does not match"):
        ops.forward_packed(x, packed16, _cfg(cfg))                 # fp32 call, bf16 pack
    cfg2 = mo.MotionConfig(64, attn_blocks=1)
    with pytest.raises(nlib.NmmError, match="does not match"):
        ops.forward_packed(x, packed32, _cfg(cfg2))


def test_unsupported_inputs_raise():
    cfg = mo.MotionConfig(64)
    m = helpers.mirror_module(cfg, mo.make_params(cfg, 1), DEV)
    with torch.no_grad():
        with pytest.raises(nlib.NmmError):
            m(torch.zeros(1, 64, 25, 2, 2, device=DEV), None, None)            # frames > max_len
        with pytest.raises(TypeError):
            m(torch.zeros(1, 64, 4, 2, 2, device=DEV, dtype=torch.float16), None, None)
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 64, 4, 2, 2, device=DEV), None, None)                 # grad mode with trainable params


def test_launch_counter_and_graph_capture():
    cfg = mo.MotionConfig(320)
    m = helpers.mirror_module(cfg, mo.make_params(cfg, 1), DEV, torch.bfloat16)
    x = mo.make_input((1, 320, 8, 16, 16), 1).to(DEV, torch.bfloat16)
    with torch.no_grad():
        y_eager = m(x, None, None)
        n0 = nb.launch_count()
        m(x, None, None)
        assert nb.launch_count() - n0 == 2                    # C = 320: GroupNorm statistics + the one-kernel module
        with nlib.options({nlib.OPT_FUSED_MODULE: 0}):
            n0 = nb.launch_count()
            m(x, None, None)
            per_call = nb.launch_count() - n0
        # gn_stats, GroupNorm-fused proj_in (16x16 positions: x is TMA-loaded by the GEMM), 2x(ln, qkv with the attention in its epilogue,
        # out), ln geglu ff_out, proj_out
        assert per_call == 1 + 1 + 2 * 3 + 3 + 1
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(x, None, None)                                  # warm-up on the side stream
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                y_graph = m(x, None, None)
        torch.cuda.current_stream().wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(y_graph, y_eager)


@pytest.mark.parametrize("C,F,H,W,B,layout", [(320, 8, 8, 8, 2, "bcfhw"), (640, 16, 4, 4, 1, "bfchw"), (64, 3, 3, 5, 2, "bcfhw"), (1280, 2, 8, 8, 1, "bfchw")])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("silu", [False, True])
def test_inflated_groupnorm_silu(C, F, H, W, B, layout, dtype, silu):
    """InflatedGroupNorm (+ SiLU) of ResnetBlock3D (resnet.py:21-29,185-186) against torch GroupNorm in fp64."""
    x = mo.make_input((B, C, F, H, W), 4, layout=layout)
    if dtype == torch.bfloat16:
        x = helpers.round_bf16(x)
    g = torch.Generator().manual_seed(C)
    w = 1 + 0.2 * torch.randn(C, generator=g)
    b = 0.2 * torch.randn(C, generator=g)
    xr = x.double().permute(0, 2, 1, 3, 4).reshape(B * F, C, H, W)                    # "b c f h w -> (b f) c h w"
    ref = torch.nn.functional.group_norm(xr, 32, w.double(), b.double(), 1e-5)
    if silu:
        ref = torch.nn.functional.silu(ref)
    ref = ref.reshape(B, F, C, H, W).permute(0, 2, 1, 3, 4)
    xd = x.to(DEV, dtype)
    assert xd.stride() == x.stride()
    y = ops.inflated_groupnorm(xd, w.to(DEV), b.to(DEV), 1e-5, silu=silu)
    assert y.is_contiguous() and y.shape == x.shape
    tol = 2e-5 if dtype == torch.float32 else 2 ** -8 * 1.01 * ref.abs().max().item()
    assert _maxabs(y, ref) <= tol


@pytest.mark.parametrize("C,F,H,W,B", [(320, 8, 8, 8, 2), (320, 16, 4, 4, 1), (640, 8, 4, 4, 2), (640, 16, 8, 8, 1), (320, 8, 4, 4, 1)])
def test_qkv_attention_fused_matches_two_kernel_path(C, F, H, W, B):
    """QKV projection with the temporal attention in its epilogue (q | k | v stay in shared memory) against
    nmm_linear + nmm_temporal_attention (same bf16 rounding of q, k, v) and against the fp64 formula."""
    cfg = mo.MotionConfig(C, max_len=32)
    nh, dh, P = cfg.heads, cfg.head_dim, H * W
    N = B * F * P
    g = torch.Generator().manual_seed(C + F)
    tok = helpers.round_bf16(torch.randn(N, C, generator=g))
    wqkv = helpers.round_bf16(torch.randn(3 * C, C, generator=g) / C ** 0.5)
    tokd, wd = tok.to(DEV, torch.bfloat16), wqkv.to(DEV, torch.bfloat16)
    fused = ops.qkv_attention(_cfg(cfg), (B, F, H, W), tokd, wd)
    qkv = ops.linear(tokd, wd, None, nlib.EPI_STORE)
    two = ops.temporal_attention(_cfg(cfg), (B, F, H, W), qkv)
    assert _maxabs(fused, two.float().cpu().double()) <= 2 ** -7 * two.float().abs().max().item()      # <= 1 bf16 ulp (P is fp32 vs hi+lo)
    z = (tok.double() @ wqkv.double().T).reshape(B, F, P, 3, nh, dh)
    s = torch.einsum("bfphd,bgphd->bphfg", z[:, :, :, 0], z[:, :, :, 1]) * dh ** -0.5
    ref = torch.einsum("bphfg,bgphd->bfphd", s.softmax(-1), z[:, :, :, 2]).reshape(N, C)
    assert _maxabs(fused, ref) <= 2e-2 * max(1.0, ref.abs().max().item())


def test_fused_and_unfused_kernel_paths_agree():
    """The three ways a C = 320 call can run: the one-kernel module (fused_module.cu), the multi-kernel pipeline with its fusions
    (GroupNorm inside proj_in, attention inside the QKV projection) and the pipeline of stand-alone kernels: same arithmetic up to
    bf16 roundings of intermediates."""
    cfg = mo.MotionConfig(320)
    params = {k: helpers.round_bf16(v) for k, v in mo.make_params(cfg, 5).items()}
    m = helpers.mirror_module(cfg, params, DEV, torch.bfloat16)
    x = helpers.round_bf16(mo.make_input((2, 320, 8, 16, 16), 6, layout="bfchw")).to(DEV, torch.bfloat16)

    def run():
        n0 = nb.launch_count()
        y = m(x, None, None).float()
        return y, nb.launch_count() - n0

    with torch.no_grad():
        m(x, None, None)                                   # first call packs the parameters (extra launches)
        y_one, n_one = run()
        with nlib.options({nlib.OPT_FUSED_MODULE: 0}):
            y_fused, n_fused = run()
            with nlib.options({nlib.OPT_GN_FUSE: 0, nlib.OPT_ATTN_FUSE: 0}):
                y_plain, n_plain = run()
    assert n_one == 2 and n_fused == 12 and n_plain == 15
    assert (y_fused - y_plain).abs().max().item() <= 2 ** -6 * y_plain.abs().max().item()
    assert (y_one - y_plain).abs().max().item() <= 2 ** -6 * y_plain.abs().max().item()
    ref = mo.forward_reference_order(params, x.float().cpu(), cfg)
    assert _maxabs(y_one, ref) <= helpers.TOL_BF16 and _maxabs(y_fused, ref) <= helpers.TOL_BF16 and _maxabs(y_plain, ref) <= helpers.TOL_BF16


@pytest.mark.parametrize("M,N,K,copy", [(16384, 640, 2560, True), (4096, 1280, 5120, False), (16384 + 128, 640, 2560, True)])
def test_linear_residual_wide_pair_tile(M, N, K, copy):
    """Shapes for which the planner picks the 256 x 320 pair tile (two N = 160 MMAs per k-step into one accumulator): ff_out at the
    C = 640 / 1280 levels.  Reference: fp64 matmul on the GPU (plumbing only)."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=DEV, generator=g)
    h = torch.randn(M, N, device=DEV, generator=g)
    ref = h.double() + A.double() @ W.double().T + b.double()
    out = ops.linear(A, W, b, nlib.EPI_RESIDUAL, h=h, want_out=copy)
    if copy:      # nmm_linear: with `out` the sum goes to `out` alone (h is only read), as in the last feed-forward of the module
        assert (out.double() - ref).abs().max().item() <= 2 ** -8 * 1.01 * ref.abs().max().item() + 1e-3
    else:
        assert (h.double() - ref).abs().max().item() <= 1e-3


# ---- fp32-grade GEMM on the tensor cores (NMM_F32X3: hi.hi + hi.lo + lo.hi bf16 products, fp32 accumulation) -------------------------
@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (130, 640, 64), (1024, 1280, 1280), (4096, 320, 1280)])
def test_linear_x3_store_and_residual(M, N, K):
    A, W, bias = _gemm_inputs(M, N, K, torch.float32)
    ref = A.double() @ W.double().T + bias.double()
    Ad, Wd, bd = A.to(DEV), W.to(DEV), bias.to(DEV)
    # two bf16 terms keep >= 16 mantissa bits of each operand: the error of a K-term dot product of N(0,1) x N(0,1/K) values has
    # rms ~4e-6 (CPU emulation: 4.4e-6 at K = 1280 and K = 5120) -> max over ~1e6 outputs ~5e-5; the module-level bar is 1e-4
    tol = 8e-5
    h = torch.full((M, N), float("nan"), device=DEV)
    assert ops.linear(Ad, Wd, bd, nlib.EPI_STORE, h=h, want_out=False, x3=True) is None
    assert _maxabs(h, ref) <= tol
    g = torch.Generator().manual_seed(11)
    h0 = torch.randn(M, N, generator=g)
    h = h0.to(DEV).clone()
    ops.linear(Ad, Wd, bd, nlib.EPI_RESIDUAL, h=h, want_out=False, x3=True)
    assert _maxabs(h, ref + h0.double()) <= tol
    # with `out`: the sum leaves as hi | lo planes (the next GEMM's A operand), h untouched; the two-term split keeps 16 mantissa bits
    h = h0.to(DEV).clone()
    out = ops.linear(Ad, Wd, bd, nlib.EPI_RESIDUAL, h=h, want_out=True, x3=True)
    assert torch.equal(h.cpu(), h0)
    assert _maxabs(out, ref + h0.double()) <= tol + 2 ** -16 * (ref + h0.double()).abs().max().item()


@pytest.mark.parametrize("M,N,K", [(256, 2560, 320), (1024, 5120, 640)])
def test_linear_x3_geglu(M, N, K):
    A, W, bias = _gemm_inputs(M, N, K, torch.float32)
    half = N // 2
    u = A.double() @ W.double().T + bias.double()
    ref = u[:, :half] * torch.nn.functional.gelu(u[:, half:])
    Wi = torch.empty_like(W); Wi[0::4] = W[0:half:2]; Wi[1::4] = W[1:half:2]; Wi[2::4] = W[half::2]; Wi[3::4] = W[half + 1::2]
    bi = torch.empty_like(bias); bi[0::4] = bias[0:half:2]; bi[1::4] = bias[1:half:2]; bi[2::4] = bias[half::2]; bi[3::4] = bias[half + 1::2]
    out = ops.linear(A.to(DEV), Wi.to(DEV), bi.to(DEV), nlib.EPI_GEGLU, x3=True)
    assert out.shape == (M, half)
    assert _maxabs(out, ref) <= 8e-5 + 2 ** -16 * ref.abs().max().item()


@pytest.mark.parametrize("C,F,H,W,B,layout", [(320, 8, 8, 8, 1, "bcfhw"), (64, 5, 3, 5, 2, "bfchw"), (640, 16, 4, 4, 1, "bfchw")])
def test_linear_x3_output_epilogue(C, F, H, W, B, layout):
    cfg = mo.MotionConfig(C)
    N = B * F * H * W
    A, Wt, bias = _gemm_inputs(N, C, C, torch.float32)
    x = mo.make_input((B, C, F, H, W), 4, layout=layout)
    y_tok = A.double() @ Wt.double().T + bias.double()
    ref = y_tok.reshape(B, F, H * W, C).permute(0, 3, 1, 2).reshape(B, C, F, H, W) + x.double()
    y = ops.linear(A.to(DEV), Wt.to(DEV), bias.to(DEV), nlib.EPI_OUTPUT, cfg=_cfg(cfg), x=x.to(DEV), x3=True)
    assert y.dtype == torch.float32 and y.shape == x.shape and y.permute(0, 2, 1, 3, 4).is_contiguous()
    assert _maxabs(y, ref) <= 8e-5
