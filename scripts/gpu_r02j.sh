# usage (under gpurun): bash scripts/gpu_r02j.sh <tag> — frame TZ search with staged windows: parity, timing vs the global-read kernels, launch list
TAG=${1:-r02j}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_golden.py -m gpu -q -x -k "tz or TZ or frame" 2>&1 | tail -8 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python microbench_tz.py --pairs 32 --steps 3 --out $O/${TAG}_tz_fs1.md 2>&1 | tail -2 | cut -c1-900
VTMME_TZ_VARIANT=global timeout 600 python microbench_tz.py --pairs 32 --steps 3 --cpu-every 50 --out $O/${TAG}_tz_fs1_global.md 2>&1 | tail -1 | cut -c1-500
timeout 600 python microbench_tz.py --pairs 32 --steps 3 --fast-search 3 --cpu-every 4 --out $O/${TAG}_tz_fs3.md 2>&1 | tail -1 | cut -c1-500
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:me_tz -c 12 --csv --log-file $O/${TAG}_tz_launches.csv python microbench_tz.py --pairs 32 --steps 1 --cpu-every 500 > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
