#!/usr/bin/env python
"""Per-kernel micro-benchmark at the shapes of one UNet step (B=2, 8 frames, 64x64 latent by default).

Times every stage of the module separately through the C ABI's per-stage entry points, with an L2 flush between
iterations, and prints achieved TFLOP/s and GB/s against the measured peaks.  Development tool; bench.py is the
contract benchmark.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neurons_b200 import lib as nlib  # noqa: E402
from neurons_b200 import ops  # noqa: E402


def timed(fn, flush, iters=5, warm=2):
    """Median device time (ms) of the library kernels `fn` launches, from the library's own CUDA events
    (host/Python launch overhead excluded), with an L2 flush before every iteration."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        nlib.profile_begin()
        fn()
        prof = nlib.profile_end()
        ts.append(sum(v["total_ms"] for v in prof.values()))
    ts.sort()
    return ts[len(ts) // 2]


def timed_chain(fns, reps):
    """Per-launch time (ms) of `reps` rounds over the variants in `fns` launched back to back in one stream: programmatic dependent
    launch active, no events or flushes between the launches; the variants use distinct buffers (together larger than L2).  The chain
    is captured in a CUDA graph so that the host (ctypes + allocator, ~20 us per call) is not what is being timed."""
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            for fn in fns:
                fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", type=int, default=0, help="N > 0: time N rounds of back-to-back launches over 4 buffer sets (in-stream cost) instead of isolated launches")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/stage_bench.json")
    ap.add_argument("--levels", default="320,640,1280", help="channel counts; 'C@side' overrides the latent side, e.g. 1280@8")
    ap.add_argument("--only", default="", help="comma-separated substrings: run only the stages whose name contains one")
    a = ap.parse_args()
    only = [t for t in a.only.split(",") if t]
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    bf = torch.bfloat16
    for lev in a.levels.split(","):
        C = int(lev.split("@")[0])
        side = int(lev.split("@")[1]) if "@" in lev else a.latent * 320 // C
        B, F = a.batch, a.frames
        P = side * side
        M = B * F * P
        cfg = ops.ModuleConfig(C)
        g = torch.Generator(device=dev).manual_seed(0)
        gw, gb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        nsets = 4 if a.chain else 1
        sets = []
        for _ in range(nsets):
            sets.append(dict(x=torch.randn(B, F, C, side, side, device=dev, dtype=bf, generator=g).permute(0, 2, 1, 3, 4),
                             act=torch.randn(M, C, device=dev, dtype=bf, generator=g), act4=torch.randn(M, 4 * C, device=dev, dtype=bf, generator=g),
                             qkv=torch.randn(M, 3 * C, device=dev, dtype=bf, generator=g), h=torch.randn(M, C, device=dev, generator=g)))
        w_cc = torch.randn(C, C, device=dev, dtype=bf, generator=g) / C ** 0.5
        w_qkv = torch.randn(3 * C, C, device=dev, dtype=bf, generator=g) / C ** 0.5
        w_1 = torch.randn(8 * C, C, device=dev, dtype=bf, generator=g) / C ** 0.5
        w_2 = torch.randn(C, 4 * C, device=dev, dtype=bf, generator=g) / (4 * C) ** 0.5
        bias = torch.randn(C, device=dev, generator=g)
        bias8 = torch.randn(8 * C, device=dev, generator=g)
        pe = torch.randn(24, C, device=dev, generator=g)
        es = 2
        # (name, factory over a buffer set, flops, bytes)
        stages = [
            ("gn_stats", lambda d: (lambda: ops.groupnorm_stats(cfg, d["x"])), 0.0, M * C * es),
            ("gn_tokens(+stats)", lambda d: (lambda: ops.groupnorm_tokens(cfg, d["x"], gw, gb)), 0.0, 3 * M * C * es),
            ("gn_stats+proj_in fused", lambda d: (lambda: ops.groupnorm_linear(cfg, d["x"], gw, gb, w_cc, bias)), 2.0 * M * C * C, M * C * (2 * es + 4)),
            ("layernorm_pe", lambda d: (lambda: ops.layernorm_pe(cfg, (B, F, side, side), d["h"], gw, gb, pe, bf)), 0.0, M * C * (4 + es)),
            ("attention", lambda d: (lambda: ops.temporal_attention(cfg, (B, F, side, side), d["qkv"])), 4.0 * M * F * C, 4 * M * C * es),
            ("qkv+attention fused", lambda d: (lambda: ops.qkv_attention(cfg, (B, F, side, side), d["act"], w_qkv)), 2.0 * M * C * 3 * C + 4.0 * M * F * C, M * C * es * 2),
            ("proj_in  C->C  store h", lambda d: (lambda: ops.linear(d["act"], w_cc, bias, nlib.EPI_STORE, h=d["h"], want_out=False)), 2.0 * M * C * C, M * C * (es + 4)),
            ("qkv      C->3C store", lambda d: (lambda: ops.linear(d["act"], w_qkv, None, nlib.EPI_STORE)), 2.0 * M * C * 3 * C, M * C * es * 4),
            ("to_out   C->C  residual", lambda d: (lambda: ops.linear(d["act"], w_cc, bias, nlib.EPI_RESIDUAL, h=d["h"], want_out=False)), 2.0 * M * C * C, M * C * (es + 8)),
            ("geglu    C->8C", lambda d: (lambda: ops.linear(d["act"], w_1, bias8, nlib.EPI_GEGLU)), 2.0 * M * C * 8 * C, M * C * es * 5),
            ("ff_out   4C->C residual+copy", lambda d: (lambda: ops.linear(d["act4"], w_2, bias, nlib.EPI_RESIDUAL, h=d["h"], want_out=True)), 2.0 * M * 4 * C * C, M * C * (4 * es + 4 + es)),
            ("proj_out C->C  nchw+x", lambda d: (lambda: ops.linear(d["act"], w_cc, bias, nlib.EPI_OUTPUT, cfg=cfg, x=d["x"])), 2.0 * M * C * C, M * C * es * 3),
        ]
        for name, make, flops, byts in stages:
            if only and not any(t in name for t in only):
                continue
            try:
                ms = timed_chain([make(d) for d in sets], a.chain) if a.chain else timed(make(sets[0]), flush)
            except nlib.NmmError as e:           # e.g. the fused QKV + attention kernel at d_h = 160
                print(f"C={C:5d} M={M:6d} {name:32s} not supported here ({e.status})", flush=True)
                continue
            rows.append(dict(C=C, side=side, M=M, stage=name, ms=ms, tflops=flops / ms / 1e9, gbps=byts / ms / 1e6,
                             frac_tensor=flops / ms / 1e9 / peaks["bf16_tflops"], frac_hbm=byts / ms / 1e6 / peaks["hbm_gbs"]))
            r = rows[-1]
            print(f"C={C:5d} M={M:6d} {name:32s} {ms*1e3:9.1f} us  {r['tflops']:8.1f} TF/s ({r['frac_tensor']*100:5.1f}%)  "
                  f"{r['gbps']:8.1f} GB/s ({r['frac_hbm']*100:5.1f}%)", flush=True)
        del sets
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
