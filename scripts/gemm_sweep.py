#!/usr/bin/env python
"""Development tool: sweep N-tile width and CTA-pair mode of the tcgen05 GEMM over the step's shapes (forces the tiling through
nmm_set_option: NMM_OPT_GEMM_BLOCK_N / NMM_OPT_GEMM_CLUSTER) and print device time per variant."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurons_b200 import lib as nlib, ops

def timed(fn, flush, iters=4):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda.synchronize()
        nlib.profile_begin(); fn(); prof = nlib.profile_end()
        ts.append(sum(v["total_ms"] for v in prof.values()))
    return sorted(ts)[len(ts) // 2]

def main():
    dev = torch.device("cuda", 0); bf = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for C, M in ((320, 65536), (640, 16384), (1280, 4096), (1280, 1024)):
        act = torch.randn(M, C, device=dev, dtype=bf, generator=g); act4 = torch.randn(M, 4 * C, device=dev, dtype=bf, generator=g)
        h = torch.randn(M, C, device=dev, generator=g); bias = torch.randn(8 * C, device=dev, generator=g)
        def W(n, k): return torch.randn(n, k, device=dev, dtype=bf, generator=g) / k ** 0.5
        wcc, wqkv, w1, w2 = W(C, C), W(3 * C, C), W(8 * C, C), W(C, 4 * C)
        gemms = {"to_out": (C, C, 32, lambda: ops.linear(act, wcc, bias[:C], nlib.EPI_RESIDUAL, h=h, want_out=False)),
                 "qkv": (3 * C, C, 32, lambda: ops.linear(act, wqkv, None, nlib.EPI_STORE)),
                 "geglu": (8 * C, C, 64, lambda: ops.linear(act, w1, bias, nlib.EPI_GEGLU)),
                 "ff_out": (C, 4 * C, 32, lambda: ops.linear(act4, w2, bias[:C], nlib.EPI_RESIDUAL, h=h, want_out=True))}
        for name, (N, K, gran, fn) in gemms.items():
            res = []
            for cg in (1, 2):
                for bn in range(256, 63, -gran):
                    if N % bn: continue
                    with nlib.options({nlib.OPT_GEMM_CLUSTER: cg, nlib.OPT_GEMM_BLOCK_N: bn}):
                        ms = timed(fn, flush)
                    res.append((ms, bn, cg))
            auto = timed(fn, flush)
            res.sort()
            fl = 2.0 * M * N * K
            print(f"C={C:4d} M={M:6d} {name:7s} auto {auto*1e3:7.1f} us | best " + "  ".join(f"bn{bn}/cg{cg}:{ms*1e3:6.1f}us({fl/ms/1e9:5.0f}TF)" for ms, bn, cg in res[:4])
                  + " | worst " + f"bn{res[-1][1]}/cg{res[-1][2]}:{res[-1][0]*1e3:.1f}", flush=True)
            rows.append(dict(C=C, M=M, gemm=name, auto_ms=auto, variants=[dict(ms=ms, bn=bn, cg=cg) for ms, bn, cg in res]))
    json.dump(rows, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
main()
