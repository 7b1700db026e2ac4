#!/bin/bash
# full GPU pass: smoke, parity tests, default bench (+reference arm), ncu launch list + full capture
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_${TAG}.json
timeout 900 python bench.py --exact 1 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_exact.json 2>&1; tail -1 gpurun_out/bench_${TAG}_exact.json | cut -c1-250
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>&1; tail -1 gpurun_out/bench_${TAG}_reference.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trace_kernel -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_dram_bytes_full_size.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_dram.log 2>&1; echo "ncu dram rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --rays 4000000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
