"""Copy the artefacts of one gpu_bench_profile.sh run (gpurun_out/*_<tag>.*) into profiles/ as r01<x>_* summaries.
usage: python scripts/save_profiles.py <tag> <prefix>      e.g.  r1e r01e"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag, pre = sys.argv[1], sys.argv[2]
shutil.copy("gpurun_out/bench_%s.json" % tag, "profiles/%s_bench.json" % pre)
shutil.copy("gpurun_out/peaks_%s.jsonl" % tag, "profiles/%s_int_peaks.jsonl" % pre)
bench = json.load(open("gpurun_out/bench_%s.json" % tag))
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    v = v / 1e3 if unit in ("us", "usecond") else v / 1e6 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("s", "second") else v
    agg[r[ix["Kernel Name"]]][0] += 1
    agg[r[ix["Kernel Name"]]][1] += v
tot = sum(v[1] for v in agg.values())
ours = [k for k in agg if "vtmme" in k and "extend_border" not in k]
tot_ours = sum(agg[k][1] for k in ours)
share = sum(agg[k][1] for k in ours if "me_tree_sad" in k) / tot_ours
with open("profiles/%s_launches_summary.md" % pre, "w") as f:
    f.write("# %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % pre)
    f.write("command: `python bench.py --steps 1 --warmup 1 --pairs-per-step 4 --pool 4 --no-cpu` "
            "(cold-cache, serialised: compare shares).\n\n")
    f.write("Share of `me_tree_sad_kernel` among the search kernels: **%.2f** here, **%.2f** live in the timed step of the full "
            "bench (`%s_bench.json`, roofline.kernel_share_of_step; the live figure also contains the host gaps of a step).\n\n"
            % (share, bench["roofline"]["kernel_share_of_step"], pre))
    f.write("| kernel | launches | total ms | share of all GPU time | share of the search kernels |\n|---|---|---|---|---|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        so = "%.1f%%" % (100 * ms / tot_ours) if k in ours else ""
        f.write("| `%s` | %d | %.3f | %.1f%% | %s |\n" % (k[:100], n, ms, 100 * ms / tot, so))
title = "%s — ncu --set full of the dominant kernel me_tree_sad_kernel<NFP=2,FPU,DY=1> (one launch = 4 pairs of 1080p, SR=64)" % pre
subprocess.check_call([sys.executable, "scripts/ncu_summary.py", "gpurun_out/prof_tree_%s.ncu-rep" % tag,
                       "profiles/%s_tree_sad_ncu.md" % pre, title], stdout=subprocess.DEVNULL)
print(open("profiles/%s_launches_summary.md" % pre).read()[:900])

# DRAM traffic of the dominant kernel for bench.py's roofline.traffic (read from this file, not pasted into the bench)
import csv as _csv, io as _io, json as _json
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_tree_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(_csv.reader(_io.StringIO(raw)))
col = {h: i for i, h in enumerate(rr[0])}
def _bytes(name):
    v, u = float(rr[2][col[name]].replace(",", "")), rr[1][col[name]]
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])
_json.dump({"kernel": "me_tree_sad_kernel", "source": "profiles/%s_tree_sad_ncu.md (ncu --set full, one launch of 4 pairs of 1080p, SR=64)" % pre,
            "pairs_per_launch": 4, "dram_bytes_read": _bytes("dram__bytes_read.sum"), "dram_bytes_write": _bytes("dram__bytes_write.sum")},
           open("profiles/ncu_traffic.json", "w"), indent=2)
print(open("profiles/ncu_traffic.json").read())
