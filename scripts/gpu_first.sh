#!/bin/bash
# first GPU pass: smoke, parity tests, bench variants, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for v in "--exact 0 --direct 0" "--exact 0 --direct 1" "--exact 1 --direct 0" "--exact 1 --direct 1"; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu $v > gpurun_out/bench_$(echo $v | tr -d ' -').log 2>&1
  echo "bench $v rc=$?"; tail -1 gpurun_out/bench_$(echo $v | tr -d ' -').log | cut -c1-400
done
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$?"; tail -1 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_r1_fast_bulk python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --rays 4000000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
