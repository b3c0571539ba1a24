#!/bin/bash
# round-2 GPU check B: parity tests, encode timing, ncu full capture of two encoder batches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests.log
python scripts/enc_time.py 1000 > gpurun_out/r2d_enc.log 2>&1
python scripts/enc_time.py 300 > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_enc_analyze|k_enc_compact' -s 18 -c 6 \
    -o gpurun_out/r2d_enc python scripts/enc_time.py 300 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_enc.log
