#!/bin/bash
# source-level capture of one full-batch k_enc_analyze launch
T=${1:-r2m}
mkdir -p gpurun_out
python scripts/enc_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_enc_analyze$' -s 6 -c 1 \
    -o gpurun_out/${T}_an python scripts/enc_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_plain.log; tail -3 gpurun_out/${T}_ncu.log
