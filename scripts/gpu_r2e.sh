#!/bin/bash
# round-2 pass e (1 GPU): full suite (tight timeout), sanitizer, fast-mode conditioning statistics
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "full pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_e.log | tail -8
bash scripts/gpu_sanitize.sh 2>&1 | tail -25
timeout 300 python tests/gpu_scripts/fast_mode_conditioning.py > gpurun_out/fast_mode_conditioning.txt 2>&1; echo "cond rc=$?"; cat gpurun_out/fast_mode_conditioning.txt | cut -c1-200
