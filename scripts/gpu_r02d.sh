# usage (under gpurun): bash scripts/gpu_r02d.sh <tag> — full GPU suite, smoke, config-5 micro-benchmark, bench (N=1)
TAG=${1:-r02d}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 1500 python microbench.py --out $O/${TAG}_microbench.md > $O/${TAG}_microbench.log 2>&1; echo "microbench rc=$?"; tail -3 $O/${TAG}_microbench.log
