# usage (under gpurun): bash scripts/gpu_r02g.sh <tag> — frame-path parity + short bench after a kernel change
TAG=${1:-r02g}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -8 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
timeout 300 python bench.py --steps 6 --warmup 3 --pool 64 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/${TAG}_bench.json"))
r=d["roofline"]
print("ms/step", d["ms_per_step"], "value", d["value"], "tree", r["kernel_ms"], "frac", r["frac"], "upper", r["other_kernels"]["me_tree_upper"]["ms"], r["other_kernels"]["me_tree_upper"]["frac"], "fracK", r["other_kernels"]["me_frac_frame"]["ms"])
print("runB", d["run_b"]["ms_per_step"], d["run_b"]["kernel_ms"])
PY
tail -3 $O/${TAG}_bench.err
