#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_custom.py -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_s2f.log; tail -3 gpurun_out/pytest_s2f.log
echo "== chain"; timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/chain_s2f.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/chain_s2f.jsonl
