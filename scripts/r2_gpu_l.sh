#!/bin/bash
# launch list of one encode call (300 streams): per-kernel durations
T=${1:-r2l}
mkdir -p gpurun_out
python scripts/enc_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -s 60 -c 40 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/enc_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_plain.log
