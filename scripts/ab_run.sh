#!/bin/bash
# on the GPU box: time encode (and decode) with every variant in build/ab/ (the in-tree library is restored at the end)
T=${1:-ab}; N=${2:-1000}
mkdir -p gpurun_out
cp flacarray_b200/libflacarray_b200.so /tmp/keep.so
for v in build/ab/*.so; do
    n=$(basename $v .so)
    cp $v flacarray_b200/libflacarray_b200.so
    echo "== $n" >> gpurun_out/${T}.log
    python scripts/enc_time.py $N 2>&1 | tail -3 >> gpurun_out/${T}.log
    if [ -n "$3" ]; then python scripts/dec_time.py $N 2>&1 | tail -3 >> gpurun_out/${T}.log; fi
done
cp /tmp/keep.so flacarray_b200/libflacarray_b200.so
cat gpurun_out/${T}.log
