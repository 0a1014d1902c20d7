#!/bin/bash
# round-2 GPU call I: column-block NW with a warp-uniform step loop: parity of the NW tests, timings, issue metrics
mkdir -p gpurun_out
L=gpurun_out/r02_i.log; : > $L
timeout 600 python -m pytest tests/test_gpu_align.py -m gpu -q --timeout 180 2>&1 | tail -4 >> $L
echo "== nw_perf" >> $L
timeout 300 python tools/nw_perf.py 24000 2>&1 | tail -1 | cut -c1-400 >> $L
echo "== c3_perf" >> $L
KG_COUNTERS=0 timeout 400 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_seed": [0-9.]*\|"ms_align": [0-9.]*\|"align_gcups": [0-9.]*' | tr '\n' ' ' >> $L
echo >> $L
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"nw_cb_kernel|nw_warp_kernel" -c 3 --csv --log-file gpurun_out/ncu_r02_i_nw.csv python tools/nw_perf.py 24000 > gpurun_out/ncu_i.log 2>&1
python - <<'PY' >> $L 2>&1
import csv
rows = list(csv.reader(open("gpurun_out/ncu_r02_i_nw.csv", errors="ignore")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
for r in rows[h + 1:]:
    if len(r) > 10:
        print(r[H.index("Kernel Name")][:40], r[H.index("Metric Name")], r[H.index("Metric Value")])
PY
cat $L
