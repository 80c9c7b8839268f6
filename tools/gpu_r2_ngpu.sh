#!/bin/bash
# N-GPU run of the bench exactly as the driver launches it:  gpurun --gpus N -- ./tools/gpu_r2_ngpu.sh N
N=${1:-8}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-200; head -c 300 gpurun_out/bench_${N}gpu.json; echo
