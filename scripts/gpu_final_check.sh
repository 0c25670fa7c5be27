# usage (under gpurun): bash scripts/gpu_final_check.sh <tag> — full GPU suite, smoke, DMVR micro-benchmark, short bench
TAG=${1:-r01j}
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/tests_$TAG.log; cat gpurun_out/tests_$TAG.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; cat gpurun_out/smoke_$TAG.log | tail -2
timeout 40 python microbench_dmvr.py --out gpurun_out/${TAG}_dmvr.md 2>&1 | tail -2
timeout 120 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json | cut -c1-600
