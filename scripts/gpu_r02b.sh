# usage (under gpurun): bash scripts/gpu_r02b.sh <tag> — the thread-per-tile refinement kernel: parity, timing against the
# warp-per-chunk kernels, ncu --set full of its five level kernels
TAG=${1:-r02c}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_mc.py -m gpu -q -x 2>&1 | tail -15 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
SMALL="--steps 4 --warmup 2 --pool 64 --no-cpu"
timeout 300 python bench.py $SMALL > $O/${TAG}_bench_tiles.json 2> $O/${TAG}_bench_tiles.err; echo "tiles rc=$?"; grep -o '"me_frac_frame": {[^}]*}' $O/${TAG}_bench_tiles.json; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_tiles.json | head -1; tail -3 $O/${TAG}_bench_tiles.err
VTMME_FRAC_VARIANT=items timeout 300 python bench.py $SMALL > $O/${TAG}_bench_items.json 2> $O/${TAG}_bench_items.err; echo "items rc=$?"; grep -o '"me_frac_frame": {[^}]*}' $O/${TAG}_bench_items.json
NCU="--steps 1 --warmup 1 --pairs-per-step 4 --pool 4 --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:me_frac_tile -c 5 -f -o $O/${TAG}_frac_tile python bench.py $NCU > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches.csv python bench.py $NCU > $O/${TAG}_ncu2.log 2>&1; echo "ncu launches rc=$?"
