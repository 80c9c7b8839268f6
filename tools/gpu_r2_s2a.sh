#!/bin/bash
# session 2 of round 2: full GPU test suite, the bench line (wall-clock timed), reference arm
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_s2a.log; tail -3 gpurun_out/pytest_s2a.log
echo "== bench"; t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s2a.json 2> gpurun_out/bench_s2a.err; echo "bench wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_s2a.err; head -c 300 gpurun_out/bench_s2a.json; echo
echo "== reference arm"; t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_s2a_ref.json 2>&1; echo "ref wall $(( $(date +%s) - t0 )) s"; head -c 300 gpurun_out/bench_s2a_ref.json; echo
