#!/bin/bash
T=${1:-r2d}
mkdir -p gpurun_out
python scripts/dec_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_dec_tile|k_dec_crc' -s 2 -c 2 \
    -o gpurun_out/${T}_dec python scripts/dec_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_plain.log; tail -2 gpurun_out/${T}_ncu.log
