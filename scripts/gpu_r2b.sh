#!/bin/bash
# round-2 second pass: new Newton path -- parity, config sweep, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_b.log
S=gpurun_out/r2b_sweep.txt; : > $S
for dt in f64 f32; do
  python scripts/sweep.py --system cooke_asph --dtype $dt default 2,1,8,2,0,0,1 2,1,8,1,0,0,1 2,2,16,1,0,0,1 2,2,8,1,0,0,1 1,1,8,2,0,0,1 >> $S 2>&1
  python scripts/sweep.py --system double_gauss --dtype $dt default >> $S 2>&1
done
python scripts/sweep.py --system cooke_asph --dtype f32 4,1,8,1,0,0,1 4,1,8,2,0,0,1 4,2,8,1,0,0,1 4,2,16,1,0,0,1 >> $S 2>&1
python scripts/sweep.py --system double_gauss --dtype f32 4,2,8,1,0,0,1 4,2,16,1,0,0,1 2,2,32,1,0,0,1 >> $S 2>&1
python scripts/sweep.py --system zoom --dtype f32 default 4,2,16,1,0,0,1 >> $S 2>&1
cat $S
for dt in f64 f32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/r2b_asph_$dt python scripts/sweep.py --system cooke_asph --dtype $dt --rays 4000000 default > gpurun_out/ncu_asph_b_$dt.log 2>&1; echo "ncu $dt rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-headline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2b.json | cut -c1-1500; tail -3 gpurun_out/bench_r2b.err
