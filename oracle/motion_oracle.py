"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the AnimateDiff motion-module forward.

This file is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it.  neurons_b200/ must not.

It restates, in plain torch-on-CPU, the algorithm of the reference hot path
(/root/reference/animatediff/models/motion_module.py, cited per function below; the inherited
diffusers-0.11.1 CrossAttention / FeedForward / GEGLU arithmetic is read from the in-tree copy
animatediff/models/motion_module_new.py:119-339,429-534 because diffusers is not vendored).

PARITY PIN: the reference holds no golden vectors / known-answer tests for this path
(SURVEY.md section 4), so the pin is the reference itself run in the build container:
tests/test_oracle_pin.py imports the unmodified reference through oracle/ref_shim.py and checks
`forward_reference_order` bit-for-bit (fp32) and `forward_token_order` to 1e-12 (fp64); the same
comparison produced the committed fixtures in tests/golden/ (oracle/gen_golden.py), which the
GPU box re-checks without the reference tree.

Two restatements:
  forward_reference_order  same op sequence as the reference (incl. its layout copies) -- used as the
                           timed CPU baseline ("port") and to be bit-identical with the reference.
  forward_token_order      the fused, copy-free formulation the CUDA kernels implement: tokens stay
                           in (b, f, p) order, attention gathers over f; returns named intermediates
                           for kernel-level parity tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch
import torch.nn.functional as TF

GN_GROUPS = 32      # motion_module.py:95,109  norm_num_groups
GN_EPS = 1e-6       # motion_module.py:109
LN_EPS = 1e-5       # nn.LayerNorm default, motion_module.py:201,207


@dataclass(frozen=True)
class MotionConfig:
    """Hyper-parameters that fix the module's shapes (motion_module.py:49-60)."""
    channels: int
    heads: int = 8                    # num_attention_heads
    layers: int = 1                   # num_transformer_block (inference-v3.yaml:10)
    attn_blocks: int = 2              # len(attention_block_types), all "Temporal_Self"
    pos_enc: bool = True              # temporal_position_encoding
    max_len: int = 24                 # temporal_position_encoding_max_len (32 for SparseCtrl)

    @property
    def head_dim(self) -> int:
        return self.channels // self.heads


def param_shapes(cfg: MotionConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys/shapes of a VanillaTemporalModule (listed from the live reference)."""
    C = cfg.channels
    s: Dict[str, Tuple[int, ...]] = {}
    t = "temporal_transformer."
    s[t + "norm.weight"] = (C,)
    s[t + "norm.bias"] = (C,)
    s[t + "proj_in.weight"] = (C, C)
    s[t + "proj_in.bias"] = (C,)
    for l in range(cfg.layers):
        b = f"{t}transformer_blocks.{l}."
        for i in range(cfg.attn_blocks):
            a = f"{b}attention_blocks.{i}."
            s[a + "to_q.weight"] = (C, C)
            s[a + "to_k.weight"] = (C, C)
            s[a + "to_v.weight"] = (C, C)
            s[a + "to_out.0.weight"] = (C, C)
            s[a + "to_out.0.bias"] = (C,)
        for i in range(cfg.attn_blocks):
            s[f"{b}norms.{i}.weight"] = (C,)
            s[f"{b}norms.{i}.bias"] = (C,)
        s[b + "ff.net.0.proj.weight"] = (8 * C, C)
        s[b + "ff.net.0.proj.bias"] = (8 * C,)
        s[b + "ff.net.2.weight"] = (C, 4 * C)
        s[b + "ff.net.2.bias"] = (C,)
        s[b + "ff_norm.weight"] = (C,)
        s[b + "ff_norm.bias"] = (C,)
    s[t + "proj_out.weight"] = (C, C)
    s[t + "proj_out.bias"] = (C,)
    return s


def make_params(cfg: MotionConfig, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, name-keyed synthetic weights (independent of module construction order).

    Linear weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the magnitude nn.Linear's default
    init gives); norm affines are perturbed away from (1, 0) so that a dropped affine is caught;
    proj_out is NOT zeroed (motion_module.py:74-75 would make the module the identity).
    """
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        is_norm = (".norm." in name) or (".norms." in name) or (".ff_norm." in name)
        if is_norm and name.endswith("weight"):
            p = 1.0 + 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        elif is_norm:
            p = 0.2 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
        else:
            fan_in = shape[1] if len(shape) == 2 else (4 * cfg.channels if name.endswith("net.2.bias") else cfg.channels)
            bound = 1.0 / math.sqrt(fan_in)
            p = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        out[name] = p.to(dtype)
    return out


def make_input(shape, seed: int, dtype=torch.float32, layout: str = "bcfhw") -> torch.Tensor:
    """Seeded N(0,1) activation of logical shape [B,C,F,H,W].

    layout "bcfhw": contiguous (the standalone benchmark tensor);
    layout "bfchw": storage order [B,F,C,H,W] viewed as [B,C,F,H,W] -- what every UNet call site
                    actually passes (producers compute on '(b f) c h w').
    Values are identical for both layouts.
    """
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype)
    if layout == "bfchw":
        x = x.permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4)
    elif layout != "bcfhw":
        raise ValueError(layout)
    return x


def positional_encoding(max_len: int, d_model: int) -> torch.Tensor:
    """Sinusoidal table [max_len, d_model], fp32 -- PositionalEncoding.__init__, motion_module.py:234-238."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _check(cfg: MotionConfig, x: torch.Tensor):
    if x.dim() != 5:                                   # motion_module.py:135
        raise AssertionError(f"Expected hidden_states to have ndim=5, but got ndim={x.dim()}.")
    B, C, F, H, W = x.shape
    if C != cfg.channels:
        raise ValueError("channel mismatch")
    if cfg.pos_enc and F > cfg.max_len:                # pe[:, :f] would broadcast-fail, motion_module.py:242
        raise ValueError("video_length exceeds temporal_position_encoding_max_len")
    return B, C, F, H, W


def forward_reference_order(params: Dict[str, torch.Tensor], x: torch.Tensor, cfg: MotionConfig) -> torch.Tensor:
    """Op-for-op restatement of VanillaTemporalModule.forward (motion_module.py:77-82 -> :134-158).

    Keeps the reference's layout copies so that its CPU cost profile (26 % copies) is
    preserved when this is timed as the CPU baseline.
    """
    B, C, F, H, W = _check(cfg, x)
    P = H * W
    nh, dh = cfg.heads, cfg.head_dim
    t = "temporal_transformer."
    pe = positional_encoding(cfg.max_len, C).to(device=x.device, dtype=x.dtype) if cfg.pos_enc else None

    h = x.permute(0, 2, 1, 3, 4).reshape(B * F, C, H, W)                    # :137  b c f h w -> (b f) c h w
    residual = h                                                           # :140
    h = TF.group_norm(h, GN_GROUPS, params[t + "norm.weight"], params[t + "norm.bias"], GN_EPS)   # :142
    h = h.permute(0, 2, 3, 1).reshape(B * F, P, C)                          # :144
    h = TF.linear(h, params[t + "proj_in.weight"], params[t + "proj_in.bias"])   # :145

    for l in range(cfg.layers):                                            # :148
        b = f"{t}transformer_blocks.{l}."
        for i in range(cfg.attn_blocks):                                   # TemporalTransformerBlock.forward :211-217
            a = f"{b}attention_blocks.{i}."
            n = TF.layer_norm(h, (C,), params[f"{b}norms.{i}.weight"], params[f"{b}norms.{i}.bias"], LN_EPS)
            # VersatileAttention.forward :270-329
            n = n.reshape(B, F, P, C).permute(0, 2, 1, 3).reshape(B * P, F, C)       # :275 (b f) d c -> (b d) f c
            if pe is not None:
                n = n + pe[None, :F]                                                 # :277-278, :242
            q = TF.linear(n, params[a + "to_q.weight"])                              # :289
            k = TF.linear(n, params[a + "to_k.weight"])                              # :297
            v = TF.linear(n, params[a + "to_v.weight"])                              # :298

            def split(z):                                                            # motion_module_new.py:181-186
                return z.reshape(B * P, F, nh, dh).permute(0, 2, 1, 3).reshape(B * P * nh, F, dh)
            q, k, v = split(q), split(k), split(v)
            s = torch.baddbmm(torch.empty(q.shape[0], F, F, dtype=q.dtype, device=q.device), q, k.transpose(-1, -2),
                              beta=0, alpha=dh ** -0.5)                               # motion_module_new.py:263-269
            p = s.softmax(dim=-1)                                                    # :277
            o = torch.bmm(p, v)                                                      # :283
            o = o.reshape(B * P, nh, F, dh).permute(0, 2, 1, 3).reshape(B * P, F, C)  # :188-193
            o = TF.linear(o, params[a + "to_out.0.weight"], params[a + "to_out.0.bias"])   # motion_module.py:321
            o = o.reshape(B, P, F, C).permute(0, 2, 1, 3).reshape(B * F, P, C)        # :327
            h = o + h                                                                # :213-217
        n = TF.layer_norm(h, (C,), params[b + "ff_norm.weight"], params[b + "ff_norm.bias"], LN_EPS)   # :219
        u = TF.linear(n, params[b + "ff.net.0.proj.weight"], params[b + "ff.net.0.proj.bias"])      # GEGLU, new:516
        val, gate = u.chunk(2, dim=-1)                                                               # new:517
        u = val * TF.gelu(gate)                                                                      # new:518 (erf gelu)
        h = TF.linear(u, params[b + "ff.net.2.weight"], params[b + "ff.net.2.bias"]) + h              # new:466, :219

    h = TF.linear(h, params[t + "proj_out.weight"], params[t + "proj_out.bias"])   # :152
    h = h.reshape(B * F, H, W, C).permute(0, 3, 1, 2).contiguous()                 # :153
    out = h + residual                                                            # :155
    return out.reshape(B, F, C, H, W).permute(0, 2, 1, 3, 4)                       # :156 (view over [B,F,C,H,W])


@dataclass
class Stages:
    """Named intermediates of forward_token_order, all in token order n = (b*F + f)*P + p."""
    gn_mean: torch.Tensor = None        # [B*F, 32]
    gn_rstd: torch.Tensor = None        # [B*F, 32]
    tokens: torch.Tensor = None         # [N, C]  GroupNorm-applied, token-major
    h0: torch.Tensor = None             # [N, C]  after proj_in
    attn_in: list = field(default_factory=list)    # LN(h)+pe per attention block
    qkv: list = field(default_factory=list)        # [N, 3C] per attention block (q | k | v)
    attn_ctx: list = field(default_factory=list)   # [N, C] softmax(QK^T)V, heads merged
    h_attn: list = field(default_factory=list)     # residual stream after each attention block
    ff_in: list = field(default_factory=list)      # LN_ff(h)
    ff_act: list = field(default_factory=list)     # [N, 4C]  value * gelu(gate)
    h_ff: list = field(default_factory=list)       # residual stream after FF
    out: torch.Tensor = None            # [B, C, F, H, W]


def forward_token_order(params: Dict[str, torch.Tensor], x: torch.Tensor, cfg: MotionConfig,
                        compute_dtype=torch.float64) -> Stages:
    """Copy-free formulation (SURVEY 8(a) restatement, re-indexed to (b, f, p) token order).

    Every Linear/LayerNorm/GEGLU is per token, GroupNorm is per (b, f, group) over (C/32 x P),
    attention is per (b, p, head) over f -- so no rearrange is needed, only a strided gather over f.
    """
    B, C, F, H, W = _check(cfg, x)
    P, N = H * W, B * F * H * W
    nh, dh = cfg.heads, cfg.head_dim
    cd = compute_dtype
    prm = {k: v.to(cd) for k, v in params.items()}
    t = "temporal_transformer."
    st = Stages()

    xf = x.to(cd).permute(0, 2, 1, 3, 4).reshape(B * F, C, P)               # [(b f), c, p]
    xg = xf.reshape(B * F, GN_GROUPS, (C // GN_GROUPS) * P)
    mean = xg.mean(dim=-1)
    var = xg.var(dim=-1, unbiased=False)
    rstd = (var + GN_EPS).rsqrt()
    st.gn_mean, st.gn_rstd = mean, rstd
    xn = ((xg - mean[..., None]) * rstd[..., None]).reshape(B * F, C, P)
    xn = xn * prm[t + "norm.weight"][None, :, None] + prm[t + "norm.bias"][None, :, None]
    tok = xn.permute(0, 2, 1).reshape(N, C)                                 # n = (bf)*P + p
    st.tokens = tok
    h = tok @ prm[t + "proj_in.weight"].T + prm[t + "proj_in.bias"]
    st.h0 = h
    pe = positional_encoding(cfg.max_len, C).to(cd) if cfg.pos_enc else None
    f_of_n = (torch.arange(N) // P) % F

    def layer_norm(z, w, b):
        mu = z.mean(dim=-1, keepdim=True)
        va = z.var(dim=-1, unbiased=False, keepdim=True)
        return (z - mu) * (va + LN_EPS).rsqrt() * w + b

    for l in range(cfg.layers):
        bk = f"{t}transformer_blocks.{l}."
        for i in range(cfg.attn_blocks):
            a = f"{bk}attention_blocks.{i}."
            n = layer_norm(h, prm[f"{bk}norms.{i}.weight"], prm[f"{bk}norms.{i}.bias"])
            if pe is not None:
                n = n + pe[f_of_n]
            st.attn_in.append(n)
            wqkv = torch.cat([prm[a + "to_q.weight"], prm[a + "to_k.weight"], prm[a + "to_v.weight"]], dim=0)
            qkv = n @ wqkv.T                                               # [N, 3C]
            st.qkv.append(qkv)
            z = qkv.reshape(B, F, P, 3, nh, dh)
            q, k, v = z[:, :, :, 0], z[:, :, :, 1], z[:, :, :, 2]          # [B, F, P, nh, dh]
            s = torch.einsum("bfphd,bgphd->bphfg", q, k) * (dh ** -0.5)
            pr = s.softmax(dim=-1)
            ctx = torch.einsum("bphfg,bgphd->bfphd", pr, v).reshape(N, C)
            st.attn_ctx.append(ctx)
            h = ctx @ prm[a + "to_out.0.weight"].T + prm[a + "to_out.0.bias"] + h
            st.h_attn.append(h)
        n = layer_norm(h, prm[bk + "ff_norm.weight"], prm[bk + "ff_norm.bias"])
        st.ff_in.append(n)
        u = n @ prm[bk + "ff.net.0.proj.weight"].T + prm[bk + "ff.net.0.proj.bias"]
        act = u[:, : 4 * C] * TF.gelu(u[:, 4 * C:])
        st.ff_act.append(act)
        h = act @ prm[bk + "ff.net.2.weight"].T + prm[bk + "ff.net.2.bias"] + h
        st.h_ff.append(h)
    y = h @ prm[t + "proj_out.weight"].T + prm[t + "proj_out.bias"]        # [N, C]
    y = y.reshape(B * F, P, C).permute(0, 2, 1) + xf                         # [(b f), c, p]
    st.out = y.reshape(B, F, C, H, W).permute(0, 2, 1, 3, 4)
    return st


def flops(cfg: MotionConfig, B: int, F: int, H: int, W: int) -> float:
    """Algorithmic FLOPs of one forward: 2*N*C^2*(2 + L*(4A + 12)) + L*A*4*N*F*C."""
    N, C = B * F * H * W, cfg.channels
    return 2.0 * N * C * C * (2 + cfg.layers * (4 * cfg.attn_blocks + 12)) + cfg.layers * cfg.attn_blocks * 4.0 * N * F * C
