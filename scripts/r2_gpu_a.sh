#!/bin/bash
# round-2 GPU check A: parity tests, encode timing, bench line, ncu launch list + full capture of the encoder kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests.log
python scripts/enc_time.py 1000 > gpurun_out/r2a_enc.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python scripts/enc_time.py 300 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_enc_analyze|k_enc_compact' -s 9 -c 3 \
    -o gpurun_out/r2a_enc python scripts/enc_time.py 300 > gpurun_out/r2a_ncu.log 2>&1
python scripts/enc_time.py 300 > gpurun_out/r2a_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2a_launches.csv python scripts/enc_time.py 300 > gpurun_out/r2a_ncu2.log 2>&1
tail -3 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_enc.log; cat gpurun_out/r2a_bench.json | cut -c1-1500
