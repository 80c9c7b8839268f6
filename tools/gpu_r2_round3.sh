#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_r2e.log; tail -3 gpurun_out/pytest_r2e.log
echo "== chain"; timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/chain_r2.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/chain_r2.jsonl
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -2 gpurun_out/bench_r2c.err; head -c 300 gpurun_out/bench_r2c.json; echo
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aux --no-e2e --pool-steps 0 > gpurun_out/ncu_launches_r2.log 2>&1
tail -1 gpurun_out/ncu_launches_r2.log | head -c 200; echo
echo "== ncu full: round kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:round_lpt --launch-skip 8 --launch-count 1 \
  -o gpurun_out/prof_round_r2 -f python tools/stream_bench.py 6 1 56832 > gpurun_out/ncu_round_r2.log 2>&1; tail -1 gpurun_out/ncu_round_r2.log | head -c 200; echo
echo "== ncu full: chain kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_chain\|ric_chain --launch-skip 2 --launch-count 2 \
  -o gpurun_out/prof_chain_r2 -f python tools/chain_bench.py 16384 2 > gpurun_out/ncu_chain_r2.log 2>&1; tail -1 gpurun_out/ncu_chain_r2.log | head -c 200; echo
