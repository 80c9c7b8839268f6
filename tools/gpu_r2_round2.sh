#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_r2d.log; tail -3 gpurun_out/pytest_r2d.log
echo "== chain"; timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/chain_r2.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/chain_r2.jsonl
echo "== variants"; timeout 600 python tools/variant_bench.py 512 2048 8192 2>&1 | tee gpurun_out/variant_r2.jsonl | cut -c1-700
echo "== mpc"; for env in "X=1" "ILQR_BURST_MAX=0" "ILQR_FWD_WPT_BELOW=100000" "ILQR_FWD_WPT_BELOW=100000 ILQR_BURST_MAX=0"; do
  echo "-- $env"; env $env timeout 300 python tools/mpc_bench.py 4096 200 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/mpc_r2.jsonl
  env $env timeout 300 python tools/mpc_bench.py 512 200 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/mpc_r2.jsonl
done
