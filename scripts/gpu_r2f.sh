#!/bin/bash
# round-2 8-GPU pass: bench --gpus 8 with C4 / C5 legs (one timed run, tight timeout)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2f_${N}gpu.json 2> gpurun_out/bench_r2f_${N}gpu.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2f_${N}gpu.json | cut -c1-7000; grep -E "Error|error|Traceback" gpurun_out/bench_r2f_${N}gpu.err | head -5
