#!/bin/bash
# r02s: ncu --set full of the rewritten table-level kernels (second launch of each = warm)
O=gpurun_out; mkdir -p $O
python scripts/dev/table_kernels_once.py > $O/r02s_plain.log 2>&1 || { cat $O/r02s_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"satd_tile_thread|interp_hor8|interp_ver8" -s 3 -c 3 -f -o $O/r02s_table python scripts/dev/table_kernels_once.py > $O/r02s_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02s_ncu.log
for i in 0 1 2; do python scripts/ncu_summary.py $O/r02s_table.ncu-rep $O/r02s_table_$i.md "r02s table kernel launch $i" $i > /dev/null 2>&1; head -30 $O/r02s_table_$i.md | grep -E "^# |time_duration|dram__bytes|dram_throughput|issue_active|registers|warps_active|pipe_alu|pipe_fma\.|pipe_lsu"; done
