#!/bin/bash
# r02t: final snapshot of the round — full GPU suite, smoke, bench + INT peaks + launch list + ncu of the dominant kernel
# (gpu_bench_profile.sh), config-5 micro-benchmark, ncu of the table-level kernels.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02t_gpu_tests.log; cat $O/r02t_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02t_smoke.log 2>&1; tail -2 $O/r02t_smoke.log
bash scripts/gpu_bench_profile.sh r2t > $O/r02t_profile.log 2>&1; grep -E "rc=" $O/r02t_profile.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02t_bench_reference.json 2> $O/r02t_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $O/r02t_bench_reference.json
timeout 1500 python microbench.py --out $O/r02t_microbench.md > $O/r02t_microbench.log 2>&1; echo "microbench rc=$?"; tail -2 $O/r02t_microbench.log
bash scripts/gpu_r02s.sh > $O/r02t_table_ncu.log 2>&1; tail -3 $O/r02t_table_ncu.log
