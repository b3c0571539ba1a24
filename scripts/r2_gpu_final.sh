#!/bin/bash
# round-2 final record: bench line (both arms), launch list of the bench command with DRAM bytes and instruction counts,
# --set full captures of the hot kernels at the full bench size, BASELINE configs 1-4 at full size
T=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1
python scripts/dec_time.py 1000 > gpurun_out/${T}_dec_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dec_tile|k_dec_crc' -s 2 -c 2 \
    -o gpurun_out/${T}_dec python scripts/dec_time.py 1000 > gpurun_out/${T}_dec_ncu.log 2>&1
python scripts/enc_time.py 1000 > gpurun_out/${T}_enc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_encode$|^k_enc_analyze$|^k_enc_compact$|^k_minmax' -s 4 -c 4 \
    -o gpurun_out/${T}_enc python scripts/enc_time.py 1000 > gpurun_out/${T}_enc_ncu.log 2>&1
python scripts/enc_time.py 300 2 > gpurun_out/${T}_enc_l2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_enc_fixed$' -s 1 -c 1 \
    -o gpurun_out/${T}_fixed python scripts/enc_time.py 300 2 > gpurun_out/${T}_fixed_ncu.log 2>&1
python scripts/run_configs.py 1 2 3 4 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err
python scripts/levels_sweep.py > gpurun_out/${T}_levels.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cut -c1-400 gpurun_out/${T}_bench.json; cat gpurun_out/${T}_dec_plain.log gpurun_out/${T}_enc_plain.log | grep -v "^it[012]"; tail -2 gpurun_out/${T}_configs.err; wc -l gpurun_out/${T}_launches.csv gpurun_out/${T}_configs.jsonl
