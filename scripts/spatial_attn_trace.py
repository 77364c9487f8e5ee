#!/usr/bin/env python
"""Per-tile clock64 timeline of CTA (0,0,0) of the tcgen05 spatial-attention kernel (trace build: python -m neurons_b200.build --trace;
NMM_LIB=libneurons_mm_trace.so python scripts/spatial_attn_trace.py [variant]).  Softmax warp 0: 0 loop top, 1 S(t) ready, 2 S in registers
(+ sfree), 3 first half of the exponentials, 4 P(t-1)V(t-1) done, 5 P stored, 6 pfull arrived.  MMA warp: 0 issue_s(t) entered, 1 K(t) landed,
2 S(t-1) released, [S(t) issued]; 3 PV(t) entered, 4 V(t) landed, 5 P(t) ready, 6 PV(t) issued."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import lib as nlib  # noqa: E402


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    dh = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lib = nlib.load()
    buf = (C.c_ulonglong * 256)()
    heads, L, images = 8, 4096 if dh == 40 else 1024, 16
    Cc = heads * dh
    qkv = torch.randn(images, L, 3 * Cc, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    nlib.set_option(nlib.OPT_SPATIAL_ATTN, variant)
    nb.spatial_attention(q, k, v, heads)
    torch.cuda.synchronize()
    lib.nmm_debug_ft_trace.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.nmm_debug_ft_trace(buf)          # allocates the device buffer
    nb.spatial_attention(q, k, v, heads)
    torch.cuda.synchronize()
    lib.nmm_debug_ft_trace(buf)
    vals = list(buf)
    t0 = min(x for x in vals if x)
    for role, name in ((0, "softmax"), (1, "mma")):
        print(f"--- {name} (cycles since the first event; variant {variant}, d_h {dh})")
        for t in range(12):
            ev = [vals[(role * 16 + t) * 8 + e] for e in range(7)]
            print(f"tile {t:2d}: " + " ".join(f"{(x - t0) if x else -1:7d}" for x in ev))


if __name__ == "__main__":
    main()
