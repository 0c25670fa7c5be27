"""Summarise an .ncu-rep (raw + source pages) into a few lines / a markdown file.  usage: ncu_summary.py rep [out.md] [title] [launch index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
sel = ["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []   # which launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
out = []
title = sys.argv[3] if len(sys.argv) > 3 else rep
out += ['# %s\n' % title, '| metric | value | unit |', '|---|---|---|']
for k in keys:
    if k in d:
        out.append('| %s | %s | %s |' % (k, d[k][1], d[k][0]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, thr, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":   # a second view of the same kernel follows: the first one is complete
        break
    if len(r) < 10:
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
    base = (m.group(2) if m else '?').split('.')[0]
    n, t = int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']])
    ops[base] += n
    thr[base] += t
    tot += n
out += ['', '## executed warp instructions by opcode\n', '| opcode | warp instr | share | avg active threads |', '|---|---|---|---|']
for k, v in ops.most_common(18):
    out.append('| %s | %d | %.2f%% | %.1f |' % (k, v, 100 * v / tot, thr[k] / max(v, 1)))
text = '\n'.join(out) + '\n'
if len(sys.argv) > 2 and sys.argv[2] != '-':
    open(sys.argv[2], 'w').write(text)
print(text)
