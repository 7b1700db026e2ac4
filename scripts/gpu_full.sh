#!/bin/bash
# full 1-GPU validation pass: smoke, parity tests, default bench, ncu launch list
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_${TAG}.log | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_${TAG}.json | cut -c1-5500; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu launches rc=$?"
