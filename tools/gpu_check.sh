#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel poisons the CUDA context) with a timeout,
# collecting logs under gpurun_out/.  Usage (on the GPU box): bash tools/gpu_check.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_canny_gpu.py tests/test_gemm_gpu.py tests/test_elementwise_gpu.py tests/test_attention_gpu.py tests/test_filter_gpu.py"}
rc=0
for f in $FILES; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --timeout 300 > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
