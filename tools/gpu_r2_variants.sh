#!/bin/bash
# Round-kernel variants on config-2 inputs (12 batches resident, one stream): warps per SM x iterations per launch.
mkdir -p gpurun_out
for cfg in "12 1" "12 2" "12 4" "12 8" "16 1" "16 4" "17 1" "17 4" "17 8"; do
  set -- $cfg
  slots=$(( 148 * ( $1 >= 16 ? 16 : 12 ) * 32 ))
  echo "== warps=$1 multi=$2 slots=$slots"
  ILQR_ROUND_WARPS=$1 ILQR_ROUND_MULTI=$2 timeout 300 python tools/stream_bench.py ${KB:-12} 1 $slots 2>&1 | tail -1 | tee -a gpurun_out/r2_variants.jsonl
done
