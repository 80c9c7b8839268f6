mkdir -p gpurun_out
timeout 300 python tools/chain_bench.py 16384 100 2>&1 | tail -1 > gpurun_out/chain_final.jsonl
timeout 600 python tools/chain_bench.py 262144 100 2>&1 | tail -1 >> gpurun_out/chain_final.jsonl; cut -c1-200 gpurun_out/chain_final.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_chain\|ric_chain --launch-skip 2 --launch-count 2 -o gpurun_out/prof_chain_final -f python tools/chain_bench.py 16384 2 > gpurun_out/ncu_chain_final.log 2>&1; tail -1 gpurun_out/ncu_chain_final.log | head -c 100; echo
