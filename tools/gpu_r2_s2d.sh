#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_chain.py -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_s2d.log; tail -3 gpurun_out/pytest_s2d.log
echo "== chain"; timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/chain_s2d.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/chain_s2d.jsonl
echo "== ncu chain"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_chain\|ric_chain --launch-skip 2 --launch-count 2 \
  -o gpurun_out/prof_chain_s2d -f python tools/chain_bench.py 16384 2 > gpurun_out/ncu_chain_s2d.log 2>&1; tail -1 gpurun_out/ncu_chain_s2d.log | head -c 200; echo
