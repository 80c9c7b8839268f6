#!/bin/bash
# round-2 evidence run: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_final.log; tail -3 gpurun_out/pytest_final.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_final.err; head -c 300 gpurun_out/bench_final.json; echo
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2>&1; head -c 200 gpurun_out/bench_final_ref.json; echo
echo "== variants / mpc"
timeout 300 python tools/variant_bench.py 512 2048 8192 > gpurun_out/variant_final.jsonl 2>&1; tail -1 gpurun_out/variant_final.jsonl | head -c 200; echo
timeout 300 python tools/mpc_bench.py 4096 500 > gpurun_out/mpc_final.jsonl 2>&1; timeout 300 python tools/mpc_bench.py 512 500 >> gpurun_out/mpc_final.jsonl 2>&1; tail -2 gpurun_out/mpc_final.jsonl | cut -c1-300
echo "== chain"; timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 > gpurun_out/chain_final.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 >> gpurun_out/chain_final.jsonl; cut -c1-300 gpurun_out/chain_final.jsonl
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aux --no-e2e --pool-steps 0 > gpurun_out/ncu_launches_final.log 2>&1
tail -1 gpurun_out/ncu_launches_final.log | head -c 200; echo
echo "== ncu full: round kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:round_lpt --launch-skip 8 --launch-count 1 \
  -o gpurun_out/prof_round_final -f python tools/stream_bench.py 6 1 56832 > gpurun_out/ncu_round_final.log 2>&1; tail -1 gpurun_out/ncu_round_final.log | head -c 200; echo
echo "== ncu full: chain kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_chain\|ric_chain --launch-skip 2 --launch-count 2 \
  -o gpurun_out/prof_chain_final -f python tools/chain_bench.py 16384 2 > gpurun_out/ncu_chain_final.log 2>&1; tail -1 gpurun_out/ncu_chain_final.log | head -c 200; echo
