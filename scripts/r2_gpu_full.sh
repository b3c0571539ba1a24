#!/bin/bash
# round-2 GPU check: parity tests, bench line (own + reference arm), launch list with instruction counts
T=${1:-r2x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python scripts/enc_time.py 300 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 40 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/enc_time.py 300 > gpurun_out/${T}_ncu.log 2>&1
tail -5 gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_bench.err; cut -c1-600 gpurun_out/${T}_bench.json
