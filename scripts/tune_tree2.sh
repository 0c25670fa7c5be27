# usage (under gpurun): bash scripts/tune_tree2.sh "<variant> ..."   variant = nfp,fpu,dy,threads,w32,bandRows
for v in $1; do
  echo "== variant $v"
  VTMME_TREE_VARIANT=$v timeout 300 python scripts/variant_parity.py 2>&1 | awk '{b+=$6} END {print "parity: bad CUs =", b}'
  VTMME_TREE_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 2 --pairs-per-step 8 --pool 8 --e2e-pool 1 --e2e-steps 1 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('RESULT $v k1_ms/pair=%.3f frac=%.3f upper=%.3f frac_k=%.3f value=%.3e' % (r['kernel_ms']/8, r['frac'], r['other_kernels_ms']['me_tree_upper']/8, r['other_kernels_ms']['me_frac_frame']/8, d['value']))
"
done
