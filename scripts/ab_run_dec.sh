#!/bin/bash
# on the GPU box: time decode with every variant in build/ab/
T=${1:-abd}; N=${2:-1000}
mkdir -p gpurun_out
cp flacarray_b200/libflacarray_b200.so /tmp/keep.so
for v in build/ab/*.so; do
    n=$(basename $v .so)
    cp $v flacarray_b200/libflacarray_b200.so
    echo "== $n" >> gpurun_out/${T}.log
    python scripts/dec_time.py $N 2>&1 | tail -3 >> gpurun_out/${T}.log
done
cp /tmp/keep.so flacarray_b200/libflacarray_b200.so
cat gpurun_out/${T}.log
