#!/bin/bash
# r02p: packed FP32 (FADD2) lanes in the tree kernel - honest mixed-pipe peaks, bare inner loop, bench per variant
mkdir -p gpurun_out
python - <<PY > gpurun_out/peaks_r02p2.jsonl 2>&1
import json
from vtm_b200.peaks import int_peak
for v in [0,6]+list(range(14,22)): print(json.dumps(int_peak(v, 1 << 16)))
PY
cat gpurun_out/peaks_r02p2.jsonl
python scripts/dev/block_bench.py 2>&1 | tee gpurun_out/r02p_block_bench.log
for v in 2 12 13 14; do
  VTMME_TREE_VARIANT=$v,1,1,256,1,0 python bench.py --steps 4 --warmup 3 > gpurun_out/r02p_bench_nfp$v.json 2> gpurun_out/r02p_bench_nfp$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02p_bench_nfp$v.json"))
print("nfp=$v", "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "tree ms %.2f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "parity", d.get("parity"), "clocks", d["clocks"], "runB", d.get("run_b", {}).get("ms_per_step"))
PY
done
