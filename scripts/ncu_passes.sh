# The ncu passes whose summaries are committed under profiles/ (B200_PROFILING.md recipe).  Run under gpurun, ONE GPU:
#   bash scripts/ncu_passes.sh && python scripts/ncu_summarise.py gpurun_out/step_metrics.csv profiles/rN_ncu_step_summary.txt profiles/rN_ncu_gemm_traffic.json
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum
timeout 600 ncu --profile-from-start off --clock-control none --csv --metrics $M --log-file gpurun_out/step_metrics.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
# launch list of the bench command itself (times only)
# (bench.py first launches ~500 torch initialisation kernels: filter on this library's kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'linear_tc|linear_simt|fused_module|gn_|layernorm_pe|temporal_attention|spatial_attention|cfg_ddim' -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clips --no-eager > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
# full sections: one C = 320 call (gn_stats + the one-kernel module) and one C = 640 call (12 kernels of the multi-kernel path)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_calls -f python scripts/profile_step.py --calls 6 --only 0,2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_calls.ncu-rep --page raw --csv > gpurun_out/prof_calls_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
