#!/bin/bash
# round-2 third pass (1 GPU): new defaults -- smoke, parity, defaults table, C3 at size, bench + reference arm, ncu evidence
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_c.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_c.log | tail -15
D=gpurun_out/r2c_defaults_all_systems.txt; : > $D
for sys in double_gauss zoom cooke cooke_asph; do for dt in f64 f32; do
  python scripts/sweep.py --system $sys --dtype $dt default >> $D 2>&1
done; done
python scripts/sweep.py --system double_gauss --exact 1 default >> $D 2>&1
python scripts/sweep.py --system cooke_asph --exact 1 default >> $D 2>&1
cat $D
python scripts/sweep.py --system cooke_asph --dtype f32 --rays 100000000 --device-rays 1 default > gpurun_out/r2_c3_1e8_f32.txt 2>&1; cat gpurun_out/r2_c3_1e8_f32.txt
python scripts/sweep.py --system cooke_asph --dtype f64 --rays 100000000 --device-rays 1 default > gpurun_out/r2_c3_1e8_f64.txt 2>&1; cat gpurun_out/r2_c3_1e8_f64.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2c.json | cut -c1-4500; tail -5 gpurun_out/bench_r2c.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/bench_r2c_reference.json 2>&1; tail -1 gpurun_out/bench_r2c_reference.json | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-headline > gpurun_out/ncu_launch_c.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trace_kernel -s 3 -c 3 --csv --log-file gpurun_out/r2c_dram_bytes_full_size.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-headline > gpurun_out/ncu_dram_c.log 2>&1; echo "ncu dram rc=$?"
