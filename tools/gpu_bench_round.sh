#!/bin/bash
# Round-end GPU pass: full test suite, default bench line, ncu launch list and one full capture of the round kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -3 gpurun_out/bench15.err; head -c 400 gpurun_out/bench15.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench15_ref.json 2>> gpurun_out/bench15.err; head -c 300 gpurun_out/bench15_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_round.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aux --no-e2e --pool-steps 0 > gpurun_out/ncu_launches_round.log 2>&1
tail -2 gpurun_out/ncu_launches_round.log | head -c 300; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:round_lpt --launch-skip 40 --launch-count 1 \
  -o gpurun_out/prof_round_r1 -f python tools/stream_bench.py 6 1 56832 > gpurun_out/ncu_round.log 2>&1
tail -2 gpurun_out/ncu_round.log
