#!/bin/bash
# round-2 first GPU pass: topology, smoke, parity tests, bench (+reference arm), aspheric baselines + ncu, FP32 budget
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | head -25 > gpurun_out/lscpu.txt; free -g >> gpurun_out/lscpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2a.json | cut -c1-3000; tail -5 gpurun_out/bench_r2a.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2a_reference.json 2>&1; tail -1 gpurun_out/bench_r2a_reference.json | cut -c1-700
timeout 300 python tests/gpu_scripts/fp32_budget.py > gpurun_out/fp32_budget.txt 2>&1; echo "fp32 rc=$?"; tail -3 gpurun_out/fp32_budget.txt
for dt in f64 f32; do
  timeout 300 python scripts/sweep.py --system cooke_asph --dtype $dt default >> gpurun_out/asph_base.txt 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/r2a_asph_$dt python scripts/sweep.py --system cooke_asph --dtype $dt --rays 4000000 default > gpurun_out/ncu_asph_$dt.log 2>&1; echo "ncu $dt rc=$?"
done
cat gpurun_out/asph_base.txt
