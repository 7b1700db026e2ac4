#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
for v in "--rpt 1 --exact 0" "--rpt 2 --exact 0" "--rpt 1 --exact 1" "--rpt 2 --exact 1"; do
  f=gpurun_out/bench_$(echo $v | tr -d ' -').log
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu $v > $f 2>&1
  echo "bench $v rc=$?"; python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
print("  value %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  clocks %s parity %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d["clocks"],d["parity_check"]))
PY
done
for r in 1 2; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_v1_fast_rpt$r python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --rays 4000000 --rpt $r > gpurun_out/ncu_full_rpt$r.log 2>&1; echo "ncu full rpt$r rc=$?"
done
ls -la gpurun_out | head -40
