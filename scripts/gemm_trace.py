#!/usr/bin/env python
"""Development tool: per-tile timeline of the tcgen05 GEMM (needs the NMM_TRACE build: python -m neurons_b200.build --trace;
run with NMM_LIB=libneurons_mm_trace.so).  Prints, for CTA 0, the cycle offsets of the MMA / TMA / epilogue events per tile."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurons_b200 import lib as nlib, ops

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    M = 65536 * 320 // C // (C // 320)
    dev = torch.device("cuda", 0)
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    act = torch.randn(M, C, device=dev, dtype=bf, generator=g)
    act4 = torch.randn(M, 4 * C, device=dev, dtype=bf, generator=g)
    h = torch.randn(M, C, device=dev, generator=g)
    bias = torch.randn(8 * C, device=dev, generator=g)
    if which == "qkvattn":
        side = int((M // 16) ** 0.5)
        cfg = ops.ModuleConfig(C)
        wq = torch.randn(3 * C, C, device=dev, dtype=bf, generator=g) / C ** 0.5
        fn = lambda: ops.qkv_attention(cfg, (2, 8, side, side), act, wq)
    W = {"qkvattn": (3 * C, C), "qkv": (3 * C, C), "geglu": (8 * C, C), "to_out": (C, C), "ff_out": (C, 4 * C)}[which]
    w = torch.randn(*W, device=dev, dtype=bf, generator=g) / W[1] ** 0.5
    fn0 = fn if which == "qkvattn" else None
    fn = {"qkvattn": fn0, "qkv": lambda: ops.linear(act, w, None, nlib.EPI_STORE),
          "geglu": lambda: ops.linear(act, w, bias, nlib.EPI_GEGLU),
          "to_out": lambda: ops.linear(act, w, bias[:C], nlib.EPI_RESIDUAL, h=h, want_out=False),
          "ff_out": lambda: ops.linear(act4, w, bias[:C], nlib.EPI_RESIDUAL, h=h, want_out=True)}[which]
    L = nlib.load()
    L.nmm_debug_trace_dump.argtypes = [ctypes.c_char_p]
    fn(); torch.cuda.synchronize()
    L.nmm_debug_trace_dump(b"/dev/null")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
    fn(); torch.cuda.synchronize()
    path = f"gpurun_out/trace_{which}_{C}.txt"
    os.makedirs("gpurun_out", exist_ok=True)
    L.nmm_debug_trace_dump(path.encode())
    rows = [[int(v) for v in ln.split()] for ln in open(path)]
    t0 = min(v for r in rows for v in r if v)
    names = ["mma:start", "mma:accfree", "mma:1stfull", "mma:issued", "tma:first", "tma:last", "epi:wait", "epi:ready", "c0:ldtm", "c0:sts", "c0:lds", "c0:stg",
             "epi:rel", "epi8:ready", "epi8:rel"]
    print(which, C, "cycles relative to first event;", " ".join(f"{n:>11s}" for n in names))
    for i, r in enumerate(rows[:14]):
        print(f"tile {i:2d}: " + " ".join(f"{(v - t0) if v else -1:11d}" for v in r[:15]))

main()
