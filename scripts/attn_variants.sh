# Development helper: attention parity tests + stage timings for the kernel variants selected by environment variables.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention or golden" 2>&1 | tail -5
for v in "NMM_ATTN_GENERIC=1" "NMM_ATTN_PB=0" "NMM_ATTN_PB=1" "NMM_ATTN_PB=3"; do
  echo "== $v"; env $v timeout 300 python scripts/stage_bench.py --only attention --levels 320,640,1280,1280@8 --out gpurun_out/sb_attn.json 2>&1 | grep attention
done
echo "== F16"; timeout 300 python scripts/stage_bench.py --only attention --frames 16 --latent 32 --levels 320,640,1280 --out gpurun_out/sb_attn.json 2>&1 | grep attention
