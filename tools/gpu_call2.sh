#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/ab_gelu_half.py 15 > gpurun_out/c2_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/c2_ab.log
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/c2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/c2_gpu_tests.log | cut -c1-300
