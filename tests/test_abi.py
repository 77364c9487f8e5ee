"""CPU: the C-ABI library loads, exports every symbol include/neurons_mm.h declares, validates shapes on the host,
and refuses to compute without an sm_100 GPU (no fallback of any kind)."""
import ctypes as C
import os
import re

import pytest
import torch

from neurons_b200 import lib as nlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "neurons_mm.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NMM_API\s+[\w\s\*]+?\b(nmm_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(built_library):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(built_library, s), f"{s} declared in neurons_mm.h but not exported"
    assert sorted(nlib.SIGNATURES.keys()) == syms, "ctypes binding and header disagree"
    assert built_library.nmm_abi_version() == 3


def _shape(**kw):
    s = nlib.Shape()
    d = dict(batch=1, channels=320, frames=8, height=4, width=4, heads=8, layers=1, attn_blocks=2, pos_enc=1, max_len=24,
             dtype=nlib.NMM_BF16, eps_gn=1e-6, eps_ln=1e-5)
    d.update(kw)
    for k, v in d.items():
        setattr(s, k, v)
    return s


def test_struct_layout_matches_header(built_library):
    # 11 int32 + 2 float + 6 int64, naturally aligned
    assert C.sizeof(nlib.Shape) == 11 * 4 + 2 * 4 + 4 + 6 * 8
    assert C.sizeof(nlib.AttnParams) == 8 * 8
    assert C.sizeof(nlib.LayerParams) == 4 * 64 + 6 * 8
    assert C.sizeof(nlib.Params) == 8 + 4 * 8 + 4 * C.sizeof(nlib.LayerParams) + 2 * 8


@pytest.mark.parametrize("kw,status", [
    (dict(), 0),
    (dict(channels=100), -1),              # not divisible by 32 GroupNorm groups
    (dict(channels=96, heads=5), -1),      # channels % heads
    (dict(frames=25), -1),                 # frames > max_len (pe[:, :f] would fail in the reference)
    (dict(frames=40, max_len=64), -2),     # beyond NMM_MAX_FRAMES
    (dict(layers=9), -2),
    (dict(attn_blocks=0), -2),
    (dict(dtype=7), -1),
    (dict(batch=0), -1),
])
def test_validate(built_library, kw, status):
    s = _shape(**kw)
    rc = built_library.nmm_validate(C.byref(s))
    assert rc == status, built_library.nmm_last_error()
    if status != 0:
        assert len(built_library.nmm_last_error()) > 0


def test_sizes(built_library):
    s = _shape(height=64, width=64)
    n = C.c_size_t()
    assert built_library.nmm_packed_params_bytes(C.byref(s), C.byref(n)) == 0
    # 2.26 M parameters in bf16 + fp32 vectors + PE tables + the tile-ordered second copy of the two q|k|v weights
    # (2 x 3 C^2) that the fused QKV + attention kernel consumes + the C = 320 one-kernel module's extras (GroupNorm-folded proj_in
    # weight, to_out tails, cumulative-bias / LayerNorm+PE vector blocks)
    assert 2 * 2_250_000 < n.value < 2 * 3_500_000
    assert built_library.nmm_workspace_bytes(C.byref(s), C.byref(n)) == 0
    # the module runs chunk by chunk over position ranges (L2-resident intermediates): the workspace holds ONE chunk
    # (tokens 2 B + residual 4 B + qkv|act 8 B + ctx 2 B per token-channel) plus the GroupNorm partial sums
    tokens = 8 * 64 * 64
    # ... plus the per-(32-row block, channel) fp32 sums of y the last kernel can emit for the next GroupNorm (8 bytes each)
    assert 4096 * 320 * (2 + 4 + 8 + 2) <= n.value <= tokens * 320 * (2 + 4 + 8 + 2) + tokens // 32 * 320 * 8 + (1 << 20)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_compute_refuses_without_gpu(built_library):
    s = _shape()
    buf = (C.c_char * 4096)()
    rc = built_library.nmm_temporal_attention(C.byref(s), buf, buf, None)
    assert rc == -5      # NMM_ERR_DEVICE: no fallback
    assert b"CUDA" in built_library.nmm_last_error()


def _spatial_shape(ctx_len=77, ctx_dim=768, **kw):
    s = nlib.SpatialShape()
    b = _shape(**kw)
    for name, _ in nlib.Shape._fields_:
        setattr(s.base, name, getattr(b, name))
    s.ctx_len, s.ctx_dim = ctx_len, ctx_dim
    return s


def test_spatial_and_decoder_sizes_and_validation(built_library):
    """Host-side checks of the SURVEY 8(f) N3 / N4 entry points (no compute): sizes, struct layouts, refusals."""
    assert C.sizeof(nlib.SpatialShape) == C.sizeof(nlib.Shape) + 8
    assert C.sizeof(nlib.SpatialLayerParams) == 20 * 8
    assert C.sizeof(nlib.SpatialParams) == 8 + 4 * 8 + 4 * C.sizeof(nlib.SpatialLayerParams) + 2 * 8
    assert C.sizeof(nlib.DecoderAttnParams) == 8 + 10 * 8
    n = C.c_size_t()
    s = _spatial_shape(height=64, width=64, batch=2)
    assert built_library.nmm_spatial_packed_params_bytes(C.byref(s), C.byref(n)) == 0
    # C^2 (2 + 3 + 1 + 1 + 1 + 8 + 4) + 2 C D weights in bf16 + fp32 vectors: ~4.5 MB at C = 320
    assert 2 * (20 * 320 * 320 + 2 * 320 * 768) < n.value < 2 * (20 * 320 * 320 + 2 * 320 * 768) + (1 << 18)
    assert built_library.nmm_spatial_workspace_bytes(C.byref(s), C.byref(n)) == 0
    tokens = 2 * 8 * 64 * 64
    assert tokens * 320 * (2 + 4 + 8 + 2) <= n.value <= tokens * 320 * (2 + 4 + 8 + 2) + tokens // 32 * 320 * 8 + (4 << 20)
    for bad, status in ((_spatial_shape(channels=256), -2),                 # head dim 32: not 40 / 80 / 160
                        (_spatial_shape(ctx_dim=100), -1),                  # ctx_dim % 8
                        (_spatial_shape(ctx_len=0), -1),
                        (_spatial_shape(dtype=nlib.NMM_F32X3, ctx_dim=96), -2)):           # 3 x bf16 mode: ctx_dim % 64
        assert built_library.nmm_spatial_workspace_bytes(C.byref(bad), C.byref(n)) == status, built_library.nmm_last_error()
    assert built_library.nmm_decoder_attn_packed_bytes(128, nlib.NMM_BF16, C.byref(n)) == 0
    assert 2 * 4 * 128 * 128 < n.value < 2 * 4 * 128 * 128 + (1 << 14)
    assert built_library.nmm_decoder_attn_packed_bytes(0, nlib.NMM_BF16, C.byref(n)) == -1
    d = _shape(channels=128, heads=1, frames=6, height=28, width=28, dtype=nlib.NMM_F32)
    assert built_library.nmm_decoder_attn_workspace_bytes(C.byref(d), C.byref(n)) == 0
    assert n.value >= 6 * 28 * 28 * 128 * 4 * 5


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_spatial_compute_refuses_without_gpu(built_library):
    buf = (C.c_char * 4096)()
    rc = built_library.nmm_spatial_attention(nlib.NMM_BF16, buf, buf, buf, buf, 960, 960, 320, 960 * 64, 960 * 64, 320 * 64, 64, 64, 8, 40, 1, 1, None)
    assert rc == -5 and b"CUDA" in built_library.nmm_last_error()
