#!/bin/bash
# ncu --set full captures of the final default kernels (one launch each, 4e6 rays)
mkdir -p gpurun_out
for spec in "double_gauss f64" "double_gauss f32" "cooke_asph f64" "cooke_asph f32" "zoom f64"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/r2k_$1_$2 python scripts/sweep.py --system $1 --dtype $2 --rays 4000000 default > gpurun_out/ncu_k_$1_$2.log 2>&1; echo "ncu $1 $2 rc=$?"
done
