#!/bin/bash
# round kernel residency experiment: 12 warps (168 registers), 14 warps (144 registers; 15 = with the parked value function), 17 = 16 warps parked
mkdir -p gpurun_out; : > gpurun_out/warps_s2g.jsonl
for rep in 1 2; do
for cfg in "12 56832" "14 66304" "15 66304" "17 75776"; do
  set -- $cfg
  ILQR_ROUND_WARPS=$1 timeout 300 python tools/stream_bench.py 20 1 $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'warps': $1, 'slots': $2, 'solves_per_s': d['solves_per_s'], 'ms_per_batch': d['ms_per_batch'], 'rounds': d['batch_iterations']}))" | tee -a gpurun_out/warps_s2g.jsonl
done; done
./tools/fp64_peak | tee gpurun_out/fp64_peak_s2g.json
