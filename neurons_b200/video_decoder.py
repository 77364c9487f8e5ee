"""Host side of SURVEY 8(f) row N4: the temporal attention + blend step of the blurry-video decoder's attention blocks
(/root/reference/model_variants/video_decoder.py:237-248 AttnUpDecoderBlock2D.forward, :394-406 UNetMidBlock2D.forward) on libneurons_mm.so,
and the batched replacement of the frame-by-frame VAE decode of pipeline_neuroclips.py:242-255.

`temporal_attention_blend(x, temp_attn, weight, time)` takes the block's own `temp_attn` module (diffusers' Attention: `group_norm`,
`to_q/to_k/to_v` with bias, `to_out[0]`, `heads`, `rescale_output_factor`) and scalar `weight` parameter and returns
`weight * x + (1 - weight) * temp_attn(...)` for x [(b t), c, h, w].  `patch_video_decoder(model)` rebinds `forward` on every block that carries
`temp_attentions` / `weights`, keeping the block's resnets, spatial attentions and upsamplers as they are.  CUDA only, inference only;
PARITY UNPINNED for this row (the diffusers class cannot be imported here: oracle/decoder_oracle.py)."""
from __future__ import annotations

import ctypes as C
import types
from typing import Optional

import torch
from torch import nn

from . import lib as _lib
from . import ops


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.NMM_F32
    if dt == torch.bfloat16:
        return _lib.NMM_BF16
    raise TypeError(f"neurons_mm supports float32 and bfloat16 activations, got {dt}")


class _DecoderEngine:
    """Packed parameters of one temp_attn + its blend weight; re-packed when any source tensor (or the weight) changes."""

    def __init__(self):
        self.key, self.packed = None, None

    def get(self, attn: nn.Module, weight: torch.Tensor, x: torch.Tensor):
        tensors = [attn.group_norm.weight, attn.group_norm.bias, attn.to_q.weight, attn.to_q.bias, attn.to_k.weight, attn.to_k.bias,
                   attn.to_v.weight, attn.to_v.bias, attn.to_out[0].weight, attn.to_out[0].bias, weight]
        key = (x.dtype, x.device) + tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self.key:
            return self.packed
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("neurons_b200: pack the decoder attention's parameters before CUDA-graph capture (run one forward first)")
        if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "norm_cross", None) is not None or attn.group_norm is None:
            raise NotImplementedError("neurons_b200: decoder temporal attention needs group_norm and no spatial_norm / norm_cross")
        if attn.to_q.bias is None or not getattr(attn, "residual_connection", True):
            raise NotImplementedError("neurons_b200: decoder temporal attention needs bias = True and residual_connection = True")
        if attn.group_norm.num_groups != 32:
            raise NotImplementedError("neurons_b200: decoder temporal attention needs 32 GroupNorm groups")
        chans = attn.to_q.weight.shape[0]
        keep = [t.detach().contiguous() for t in tensors[:-1]]
        if any(t.device != x.device for t in keep):
            raise RuntimeError("neurons_b200: decoder attention parameters and input are on different devices")
        if len({t.dtype for t in keep}) != 1:
            raise TypeError("neurons_b200: mixed parameter dtypes in the decoder attention")
        p = _lib.DecoderAttnParams()
        p.dtype = _dtype_code(keep[0].dtype)
        for name, t in zip(("gn_w", "gn_b", "to_q_w", "to_q_b", "to_k_w", "to_k_b", "to_v_w", "to_v_b", "to_out_w", "to_out_b"), keep):
            setattr(p, name, t.data_ptr())
        lib = _lib.load()
        n = C.c_size_t()
        _lib.check(lib.nmm_decoder_attn_packed_bytes(chans, _dtype_code(x.dtype), C.byref(n)))
        packed = torch.empty(n.value, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.nmm_decoder_attn_pack(chans, _dtype_code(x.dtype), C.byref(p), float(weight.detach().float().reshape(-1)[0]),
                                                 float(getattr(attn, "rescale_output_factor", 1.0)), packed.data_ptr(), n.value,
                                                 ops._stream_ptr(x.device)))
        torch.cuda.current_stream(x.device).synchronize()
        del keep
        self.key, self.packed = key, packed
        return packed


def temporal_attention_blend(x: torch.Tensor, temp_attn: nn.Module, weight: torch.Tensor, time: int) -> torch.Tensor:
    """video_decoder.py:241-248 in one call.  x: [(b t), c, h, w] contiguous (what the block's resnet / spatial attention hand over)."""
    if x.dim() != 4 or x.shape[0] % time != 0:
        raise ValueError(f"expected [(b t), c, h, w] with t = {time}, got {tuple(x.shape)}")
    ops._require_cuda(x, "hidden_states")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in temp_attn.parameters())):
        raise RuntimeError("neurons_b200: the decoder temporal attention op is inference-only; call it under torch.no_grad()")
    x = x.contiguous()
    eng = temp_attn.__dict__.get("_nmm_engine")
    if eng is None:
        eng = temp_attn.__dict__["_nmm_engine"] = _DecoderEngine()
    packed = eng.get(temp_attn, weight, x)
    bt, c, h, w = x.shape
    b = bt // time
    s = _lib.Shape()
    s.batch, s.channels, s.frames, s.height, s.width = b, c, time, h, w
    s.heads, s.layers, s.attn_blocks, s.pos_enc, s.max_len = int(temp_attn.heads), 1, 1, 0, 0
    s.dtype, s.eps_gn, s.eps_ln, s.ln_fold = _dtype_code(x.dtype), float(temp_attn.group_norm.eps), ops.LN_EPS, 0
    s.x_stride_b, s.x_stride_c, s.x_stride_f = time * c * h * w, h * w, c * h * w          # [(b t), c, h, w] = [B, F, C, H, W] storage
    s.y_stride_b, s.y_stride_c, s.y_stride_f = s.x_stride_b, s.x_stride_c, s.x_stride_f
    y = torch.empty_like(x)
    lib = _lib.load()
    n = C.c_size_t()
    _lib.check(lib.nmm_decoder_attn_workspace_bytes(C.byref(s), C.byref(n)))
    ws, ws_ptr = ops._aligned_ws(n.value, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.nmm_decoder_temporal_attention(C.byref(s), x.data_ptr(), y.data_ptr(), packed.data_ptr(), packed.numel(), ws_ptr, n.value,
                                                      ops._stream_ptr(x.device)))
    return y


def _up_block_forward(self, hidden_states, temb=None, scale: float = 1.0, time: int = 1):
    """AttnUpDecoderBlock2D.forward, video_decoder.py:233-255, with :241-248 replaced by the fused op."""
    for attn, temp_attn, resnet, weight in zip(self.attentions, self.temp_attentions, self.resnets, self.weights):
        hidden_states = resnet(hidden_states, temb)
        if attn is not None:
            hidden_states = attn(hidden_states, temb=temb, scale=scale)
            hidden_states = temporal_attention_blend(hidden_states, temp_attn, weight, time)
    if self.upsamplers is not None:
        for upsampler in self.upsamplers:
            hidden_states = upsampler(hidden_states)
    return hidden_states


def _mid_block_forward(self, hidden_states, temb=None, time: int = 1):
    """UNetMidBlock2D.forward, video_decoder.py:394-409, with :398-406 replaced by the fused op."""
    hidden_states = self.resnets[0](hidden_states, temb)
    for attn, temp_attn, resnet, weight in zip(self.attentions, self.temp_attentions, self.resnets[1:], self.weights):
        if attn is not None:
            hidden_states = attn(hidden_states, temb=temb)
            hidden_states = temporal_attention_blend(hidden_states, temp_attn, weight, time)
        hidden_states = resnet(hidden_states, temb)
    return hidden_states


def patch_video_decoder(model: nn.Module) -> int:
    """Rebind `forward` on every AttnUpDecoderBlock2D / UNetMidBlock2D of a DecoderVideo (blocks that carry `temp_attentions` and
    `weights`).  Returns the number of blocks patched."""
    n = 0
    for m in model.modules():
        if hasattr(m, "temp_attentions") and hasattr(m, "weights") and hasattr(m, "resnets") and hasattr(m, "attentions"):
            m.forward = types.MethodType(_up_block_forward if hasattr(m, "upsamplers") else _mid_block_forward, m)
            n += 1
    return n


def decode_latents_batched(vae_decode, latents: torch.Tensor, chunk: Optional[int] = None) -> torch.Tensor:
    """pipeline_neuroclips.py:242-255 (`decode_latents`) decodes a clip one frame at a time (16 serial VAE launches per clip); this is the
    same arithmetic in `chunk`-frame batches (default: all frames at once).  latents: [b, c, f, h, w] as the sampler returns them (the
    1 / 0.18215 scaling of :244 is applied here); vae_decode: frames [n, c, h, w] -> images [n, 3, H, W] (the reference:
    `lambda z: vae.decode(z).sample`).  Returns [b, 3, f, H, W] in [0, 1], float32, on the latents' device (the reference's `.cpu().numpy()`
    is the caller's)."""
    b, c, f, h, w = latents.shape
    frames = (1 / 0.18215 * latents).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)      # 'b c f h w -> (b f) c h w'
    step = chunk or frames.shape[0]
    video = torch.cat([vae_decode(frames[i:i + step]) for i in range(0, frames.shape[0], step)])
    video = video.reshape(b, f, *video.shape[1:]).permute(0, 2, 1, 3, 4)                 # '(b f) c h w -> b c f h w'
    return (video / 2 + 0.5).clamp(0, 1).float()
