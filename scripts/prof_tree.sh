# usage: bash scripts/prof_tree.sh "<variant> ..."  : parity tests with the default variant, then one ncu --set full capture per variant
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
SMALL="--steps 1 --warmup 1 --pairs-per-step 4 --pool 4 --no-cpu"
for v in $1; do
  tag=$(echo $v | tr ',' '_')
  VTMME_TREE_VARIANT=$v timeout 600 python bench.py $SMALL > gpurun_out/plain_$tag.log 2>&1 &&
  VTMME_TREE_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:me_tree_sad -s 1 -c 1 -f -o gpurun_out/prof_tree_$tag python bench.py $SMALL > gpurun_out/ncu_$tag.log 2>&1
  echo "variant $v rc=$?"; grep -o '"kernel_ms": [0-9.]*' gpurun_out/plain_$tag.log
done
