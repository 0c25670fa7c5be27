#!/bin/bash
# r02q: table-level kernels (pair-wise interpolation, thread-per-tile SATD): targeted tests, then the config-5 micro-benchmark
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_table_batch.py tests/test_gpu_golden.py tests/test_gpu_parity.py -q -x -k "table or satd or interp or filter or dist or golden" 2>&1 | tail -15 > $O/r02q_table_tests.log; cat $O/r02q_table_tests.log
timeout 1500 python microbench.py --out $O/r02q_microbench.md > $O/r02q_microbench.log 2>&1; echo "microbench rc=$?"; tail -3 $O/r02q_microbench.log
grep -E "\| (8x8|16x16|32x32|64x64|128x128|64x16|16x64) \| uniform" $O/r02q_microbench.md | grep -E "SATD|luma8|chroma4"
