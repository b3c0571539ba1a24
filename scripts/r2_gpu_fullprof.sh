#!/bin/bash
# full-size (1000 streams) --set full captures: one launch each of the tile decoder, the CRC pass and the encoder kernels
T=${1:-r2q}
mkdir -p gpurun_out
python scripts/dec_time.py 1000 > gpurun_out/${T}_dec_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dec_tile|k_dec_crc' -s 2 -c 2 \
    -o gpurun_out/${T}_dec python scripts/dec_time.py 1000 > gpurun_out/${T}_dec_ncu.log 2>&1
python scripts/enc_time.py 1000 > gpurun_out/${T}_enc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_encode$|^k_enc_analyze$|^k_enc_compact$' -s 6 -c 3 \
    -o gpurun_out/${T}_enc python scripts/enc_time.py 1000 > gpurun_out/${T}_enc_ncu.log 2>&1
cat gpurun_out/${T}_dec_plain.log gpurun_out/${T}_enc_plain.log; tail -2 gpurun_out/${T}_dec_ncu.log gpurun_out/${T}_enc_ncu.log
