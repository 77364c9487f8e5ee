# Development probe: where does the GroupNorm-fused proj_in spend its time?  NMM_GEMM_DEBUG bits: 1 no epilogue, 2 no TMA,
# 4 converter skips the arithmetic, 8 no proxy fence, 16 K-major descriptors (results invalid in every mode).
for d in 0 4; do
  echo "== NMM_GEMM_DEBUG=$d"; NMM_GEMM_DEBUG=$d timeout 300 python scripts/stage_bench.py --levels 320,640,1280,1280@8 --only "proj_in fused" --out gpurun_out/sb_gn.json 2>&1 | grep "C="
done
