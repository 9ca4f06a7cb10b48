#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel poisons the CUDA context of that process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in ${@:-tests/test_gpu_igemm.py tests/test_gpu_attention.py tests/test_gpu_norm_elem.py}; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --no-header -rA 2>&1 | tail -150 > gpurun_out/$n.log
  echo "== $n exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/$n.log
done
