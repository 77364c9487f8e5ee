#!/bin/bash
# Ordered GPU parity run: safe kernels first, tcgen05 last, every group under its own timeout so a hung kernel
# cannot eat the box.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, pytest -k expression
  echo "=== $1 ===" | tee -a gpurun_out/check.log
  timeout "$2" python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "$3" 2>&1 | tail -${4:-40} | tee -a gpurun_out/check.log
  echo "exit=$?" | tee -a gpurun_out/check.log
}
run norms_attention 600 "groupnorm or layernorm or attention"
run linear_fp32 600 "linear and float32"
run module_fp32 900 "golden_fp32 or (config1 and float32) or (properties and float32) or lora or unsupported"
run linear_bf16_tcgen05 300 "linear and bfloat16" 80
run module_bf16 900 "golden_bf16 or bf16_from or (config1 and bfloat16) or (properties and bfloat16) or launch_counter" 80
