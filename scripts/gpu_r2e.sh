#!/bin/bash
# round-2 pass e (1 GPU): generator / batched front end / dropin tests (tight timeouts), then the full suite, sanitizer
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_aim.py -q -x > gpurun_out/pytest_gpu_e1.log 2>&1; echo "aim pytest rc=$?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_gpu_e1.log | tail -8
timeout 300 python -m pytest tests/test_gpu_dropin_reference.py -q -x > gpurun_out/pytest_gpu_e2.log 2>&1; echo "dropin pytest rc=$?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_gpu_e2.log | tail -8
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "full pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_e.log | tail -8
sed -i 's/timeout 900 compute-sanitizer/timeout 400 compute-sanitizer/' scripts/gpu_sanitize.sh
bash scripts/gpu_sanitize.sh 2>&1 | tail -25
