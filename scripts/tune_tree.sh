# usage (under gpurun): bash scripts/tune_tree.sh "<variant> <variant> ..."   variant = nfp,fpu,dy,threads
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in $1; do
  echo "== variant $v"
  VTMME_TREE_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 1 --pairs-per-step 8 --pool 8 --e2e-pool 1 --e2e-steps 1 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('RESULT $v k1_ms/pair=%.3f frac=%.3f upper=%.3f frac_k=%.3f value=%.3e clocks=%s' % (r['kernel_ms']/8, r['frac'], r['other_kernels_ms']['me_tree_upper']/8, r['other_kernels_ms']['me_frac_frame']/8, d['value'], d['clocks']))
    else: print(line)
"
done
