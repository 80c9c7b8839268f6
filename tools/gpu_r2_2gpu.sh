#!/bin/bash
# 2-GPU run of the bench exactly as the driver launches it
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/bench_2gpu.err; head -c 400 gpurun_out/bench_2gpu.json; echo
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>&1
echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/bench_2gpu_ref.json | head -c 300; echo
