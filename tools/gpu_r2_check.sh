#!/bin/bash
# quick round-end check: full GPU test suite, smoke, bench line (both arms)
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_check.log; tail -3 gpurun_out/pytest_check.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_check.err; head -c 300 gpurun_out/bench_check.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_check_ref.json 2>&1; head -c 200 gpurun_out/bench_check_ref.json; echo
