#!/bin/bash
# launch list with instruction counts + full captures of the encoder kernels (one batch) for the source pages
T=${1:-r2p}
mkdir -p gpurun_out
python scripts/enc_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 30 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/enc_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
python scripts/enc_time.py 300 > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_encode$|^k_enc_analyze$|^k_enc_compact$' -s 6 -c 3 \
    -o gpurun_out/${T}_enc python scripts/enc_time.py 300 > gpurun_out/${T}_ncu2.log 2>&1
cat gpurun_out/${T}_plain.log; tail -3 gpurun_out/${T}_ncu2.log
