#!/bin/bash
# configs[3] (7-DoF chain) backward pass: split (analytic) vs dual-number kernel, B = 16,384 and full size
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_r2c.log
tail -3 gpurun_out/pytest_r2c.log
./tools/fp64_peak > gpurun_out/fp64_peak_r2.json 2>&1; cat gpurun_out/fp64_peak_r2.json
for an in 1 0; do
  echo "== chain_bench analytic=$an B=16384"
  ILQR_CHAIN_ANALYTIC=$an timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/chain_r2.jsonl
done
echo "== chain_bench analytic=1 B=262144"
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/chain_r2.jsonl
