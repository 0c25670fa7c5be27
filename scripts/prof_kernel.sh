# usage: bash scripts/prof_kernel.sh <kernel-regex> <tag> [count]
set -x
SMALL="--steps 1 --warmup 1 --pairs-per-step 4 --pool 4 --no-cpu"
timeout 600 python bench.py $SMALL > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${4:-0} -c ${3:-1} -f -o gpurun_out/prof_$2 python bench.py $SMALL > gpurun_out/ncu_$2.log 2>&1
echo "rc=$?"
