#!/bin/bash
# 8-GPU run of the bench exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8_r2.txt 2>&1
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/bench_8gpu.err | cut -c1-300; head -c 400 gpurun_out/bench_8gpu.json; echo
