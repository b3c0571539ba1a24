#!/bin/bash
# round-2 GPU check E: parity tests, encode / decode timing, bench line, ncu full capture of encoder batch + tile decoder
T=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
python scripts/enc_time.py 1000 > gpurun_out/${T}_enc.log 2>&1
python scripts/dec_time.py 1000 > gpurun_out/${T}_dec.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python scripts/enc_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_enc_analyze|k_enc_compact' -s 21 -c 3 \
    -o gpurun_out/${T}_enc python scripts/enc_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
python scripts/dec_time.py 300 > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_dec_tile|k_dec_crc' -s 2 -c 2 \
    -o gpurun_out/${T}_dec python scripts/dec_time.py 300 > gpurun_out/${T}_ncu2.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_enc.log gpurun_out/${T}_dec.log; cut -c1-900 gpurun_out/${T}_bench.json
