# usage (under gpurun): bash scripts/gpu_bench_profile.sh <tag>
TAG=${1:-r1}
set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/peaks_$TAG.jsonl 2>&1
import json
from vtm_b200.peaks import all_peaks
for r in all_peaks(1 << 16): print(json.dumps(r))
PY
cat gpurun_out/peaks_$TAG.jsonl
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
SMALL="--steps 1 --warmup 1 --pairs-per-step 4 --pool 4 --no-cpu"
timeout 600 python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch-list rc=$?"
timeout 600 python bench.py $SMALL > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:me_tree_sad -s 1 -c 1 -f -o gpurun_out/prof_tree_$TAG python bench.py $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
