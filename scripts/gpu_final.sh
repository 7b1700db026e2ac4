#!/bin/bash
# final pass of the round (1 GPU): smoke, all GPU tests, defaults table, C3 at size, bench
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_m.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_m.log
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_m.log | tail -6
D=gpurun_out/r2m_defaults_all_systems.txt; : > $D
for sys in double_gauss zoom cooke cooke_asph; do for dt in f64 f32; do python scripts/sweep.py --system $sys --dtype $dt default 2>&1 | grep tune >> $D; done; done
python scripts/sweep.py --system double_gauss --exact 1 default 2>&1 | grep tune >> $D
python scripts/sweep.py --system cooke_asph --exact 1 default 2>&1 | grep tune >> $D
cat $D | cut -c60-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo "bench rc=$?"; python - <<EOF
import json
d=json.loads(open("gpurun_out/bench_r2m.json").read().strip().splitlines()[-1])
print(d["parity_ok"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["spot_consumer"]["value"])
print(d["headline"]["kernel_ms"], d["headline"]["frac"], d["c3"]["kernel_ms"], d["c3"]["frac"], d["c3"]["parity"])
print(d["cpu_baseline"])
EOF
