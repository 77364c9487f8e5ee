#!/usr/bin/env python
"""Per-phase timeline of the fused module kernel (CTA 0, first tile) from the -DNMM_TRACE build:
    python -m neurons_b200.build --trace && NMM_LIB=libneurons_mm_trace.so python scripts/fused_trace.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurons_b200 as nb  # noqa: E402
from neurons_b200 import lib as nlib  # noqa: E402

B, F, side = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (2, 8, 16)))
EMIT = len(sys.argv) >= 5 and sys.argv[4] == "emit"      # also emit the GroupNorm sums of y (N1)
dev = torch.device("cuda", 0)
kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
          temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
with torch.no_grad():
    with torch.device(dev):
        m = nb.get_motion_module(320, "Vanilla", kw).to(torch.bfloat16).eval()
    x = torch.randn(B, F, 320, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4)
    m(x, None, None)
    torch.cuda.synchronize()
    lib = nlib.load()
    lib.nmm_debug_fm_trace_dump(b"/dev/null")
    if EMIT:
        from neurons_b200 import ops
        eng = m.__dict__["_nmm_engine"]
        ops.forward_packed(x, eng.packed, eng.cfg, y_sums=torch.empty(B * F * 32, 2, dtype=torch.float64, device=dev))
    else:
        m(x, None, None)
    torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", "fm_trace.txt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    assert lib.nmm_debug_fm_trace_dump(path.encode()) == 0
t = {}
for line in open(path):
    i, v = line.split()
    if int(v):
        t[int(i)] = int(v)
t0 = t[0]
names = {0: "tile start", 1: "x staged", 2: "tokens published", 3: "proj_in done", 44: "LN_ff published", 65: "ff done", 66: "bf16 h published",
         67: "proj_out done", 68: "y stored", 100: "mma: tokens ready", 101: "mma: proj_in issued", 198: "mma: bf16 h ready", 199: "mma: proj_out issued"}
for i in range(2):
    names[4 + 20 * i] = f"attn{i}: LN published"
    names[4 + 20 * i + 17] = f"attn{i}: to_out done"
    names[102 + 20 * i + 15] = f"mma attn{i}: last to_out issued"
    for hp in range(4):
        names[4 + 20 * i + 1 + 4 * hp] = f"attn{i} pair{hp}: q|k|v dumped"
        names[4 + 20 * i + 2 + 4 * hp] = f"attn{i} pair{hp}: attention done"
        names[4 + 20 * i + 3 + 4 * hp] = f"attn{i} pair{hp}: ctx copied"
        names[102 + 20 * i + 12 + hp] = f"mma attn{i}: to_out pair{hp} issued"
        for s in range(3):
            names[102 + 20 * i + 3 * hp + s] = f"mma attn{i}: unit ({hp},{s}) issued"
for hp in range(4):
    for s_ in range(3):
        names[220 + 3 * hp + s_] = f"prod attn0: unit ({hp},{s_}) fills issued"
for j in range(20):
    names[232 + j] = f"prod: G_{j} fills issued"
    names[45 + j] = f"ff chunk {j}: act published"
    names[150 + j] = f"mma: G_{j} issued"
    names[175 + j] = f"mma: F_{j} issued"
names.update({240: "chunk10: start (waiting for the accumulator)", 241: "chunk10: accumulator ready", 242: "chunk10: loaded from TMEM", 243: "chunk10: GELUs done",
              244: "chunk10: activation buffer free"})
prev = {0: t0, 1: t0, 2: t0}
for k in sorted(t, key=lambda k: t[k]):
    side_ = 0 if (k < 100 or k >= 240) else (1 if k < 220 else 2)
    print(f"{t[k] - t0:9d}  (+{t[k] - prev[side_]:7d})  {('EPI', 'MMA', 'PRD')[side_]}  {names.get(k, k)}")
    prev[side_] = t[k]
