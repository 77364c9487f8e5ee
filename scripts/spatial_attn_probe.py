#!/usr/bin/env python
"""Development probe: the spatial-attention kernels against an fp64 softmax attention, for the run-time variants of NMM_OPT_SPATIAL_ATTN
(0 = tcgen05 kernel as planned: two query tiles per CTA; 1 = mma.sync kernel; 23 = one query tile per CTA, two CTAs per SM; 2 = that with two
softmax threads per query row; 10..14 = polynomial-ex2 share 0 / 2 / 3 / 4 / 6 of 8; 15 = no K / V traffic (timing only); 16 = cp.async loader;
20 / 21 / 22 = two query tiles with 2 / 3 / 4 K / V stages), plus timing at the 64 x 64 (d_h 40) and 32 x 32 (d_h 80) levels.   python scripts/spatial_attn_probe.py 1,0,2"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import lib as nlib  # noqa: E402


def ref64(q, k, v, heads):
    I, L, C = q.shape
    dh = C // heads
    qq, kk, vv = (t.double().view(I, -1, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    p = torch.softmax(qq @ kk.transpose(-1, -2) * dh ** -0.5, dim=-1)
    return (p @ vv).permute(0, 2, 1, 3).reshape(I, L, C)


def main():
    variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 0, 2]
    for dh, L, images, scale in ((40, 256, 1, 1.0), (40, 1024, 2, 1.0), (80, 384, 2, 1.0), (40, 300, 1, 4.0), (80, 1024, 1, 3.0)):
        heads = 8
        C = heads * dh
        g = torch.Generator().manual_seed(dh + L)
        qkv = (torch.randn(images, L, 3 * C, generator=g) * scale).to(torch.bfloat16).cuda()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        ref = ref64(q.cpu(), k.cpu(), v.cpu(), heads)
        for var in variants:
            nlib.set_option(nlib.OPT_SPATIAL_ATTN, var)
            o = nb.spatial_attention(q, k, v, heads)
            torch.cuda.synchronize()
            err = (o.double().cpu() - ref).abs().max().item()
            print(f"dh={dh} L={L} images={images} scale={scale} variant={var}: max-abs err {err:.3e}", flush=True)
    # timing at the 64 x 64 level (C = 320, 16 images) and the 32 x 32 level (C = 640)
    for dh, L, images in ((40, 4096, 16), (80, 1024, 16)):
        heads = 8
        C = heads * dh
        qkv = torch.randn(images, L, 3 * C, device="cuda", dtype=torch.bfloat16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        for var in variants:
            nlib.set_option(nlib.OPT_SPATIAL_ATTN, var)
            for _ in range(2):
                nb.spatial_attention(q, k, v, heads)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                nb.spatial_attention(q, k, v, heads)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 5
            print(f"dh={dh} L={L} images={images} variant={var}: {us:8.1f} us  {4.0 * images * L * L * C / us / 1e6:7.1f} TF/s", flush=True)
    nlib.set_option(nlib.OPT_SPATIAL_ATTN, 0)


if __name__ == "__main__":
    main()
