#!/bin/bash
# round-2 multi-GPU pass (N GPUs of one box): a few tests, bench --gpus N with C4/C5 legs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_systems.py tests/test_gpu_aim.py -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_d.log | tail -8
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2d_${N}gpu.json 2> gpurun_out/bench_r2d_${N}gpu.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2d_${N}gpu.json | cut -c1-7000; grep -E "Error|error|Traceback" gpurun_out/bench_r2d_${N}gpu.err | head -5
