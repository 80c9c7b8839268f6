#!/bin/bash
# GPU check of the fused streaming rounds: parity tests, then throughput (stream solve, streamer), optional ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream" 2>&1 | tail -5
for w in 16; do for sh in 0 1; do
  echo "== parity warps=$w shift=$sh"
  ILQR_ROUND_WARPS=$w ILQR_ROUND_SHIFT=$sh timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k stream 2>&1 | tail -2
done; done
echo "== stream solve (device, one call)"
timeout 300 python tools/stream_bench.py 12 1 56832 2>&1 | tail -1 | tee -a gpurun_out/round_bench.jsonl
echo "== streamer"
timeout 600 python tools/streamer_bench.py ${KB:-24} 8 56832 2>&1 | tail -1 | tee -a gpurun_out/round_bench.jsonl
if [ -n "$NCU" ]; then
ILQR_ROUND_WARPS=12 ILQR_ROUND_SHIFT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:round_lpt --launch-skip 40 --launch-count 1 \
  -o gpurun_out/prof_round_r1 -f python tools/stream_bench.py 2 1 56832 > gpurun_out/ncu_round.log 2>&1
tail -3 gpurun_out/ncu_round.log
fi
