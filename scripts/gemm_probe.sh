#!/bin/bash
# timing experiments on the GEMM kernel (results numerically invalid when NMM_GEMM_DEBUG != 0)
for cl in 1; do for dbg in 0 2 10 18 26; do
  echo "== cluster=$cl debug=$dbg"
  NMM_GEMM_CLUSTER=$cl NMM_GEMM_DEBUG=$dbg timeout 120 python scripts/stage_bench.py --levels 320 --out gpurun_out/probe.json 2>&1 | grep -E "qkv|geglu|to_out"
done; done
