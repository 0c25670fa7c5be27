#!/bin/bash
# r02w: last GPU session of the round — full GPU suite, smoke and a bench line on the final code, then config 3 (2160p, SearchRange
# 128) with 8 pictures against the CPU golden
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02w_gpu_tests.log; cat $O/r02w_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02w_smoke.log 2>&1; tail -2 $O/r02w_smoke.log
timeout 600 python bench.py > $O/r02w_bench.json 2> $O/r02w_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r02w_bench.json
TMO=${TMO:-3000} bash scripts/gpu_r02u.sh "3 8"
