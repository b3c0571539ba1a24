#!/bin/bash
# quick GPU check: parity tests, encode/decode timing, optional sanitizer pass
T=${1:-r2q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
python scripts/enc_time.py 1000 > gpurun_out/${T}_enc.log 2>&1
python scripts/dec_time.py 1000 > gpurun_out/${T}_dec.log 2>&1
if [ -n "$2" ]; then
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/${T}_race.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_race.log
  timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/${T}_mem.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_mem.log
  tail -4 gpurun_out/${T}_race.log; tail -4 gpurun_out/${T}_mem.log
fi
tail -4 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_enc.log; tail -2 gpurun_out/${T}_dec.log
