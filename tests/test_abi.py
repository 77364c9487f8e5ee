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
