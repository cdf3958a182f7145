#!/bin/bash
# Round-2 lab run Q: cluster kernel, columns and rows with the most pushes first.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02q
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 600 python -m pytest tests/test_gpu_k16384.py -q -x > $OUT/pytest_k16384.txt 2>&1; echo "k16384 rc=$?" >> $OUT/log.txt; tail -3 $OUT/pytest_k16384.txt
for spec in "11 i8 8192 1.8" "10 i8 8192 2.6" "9 i8 8192 3.6"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm_cluster -s 1 -c 1 -f -o $OUT/r02_tm32768_cluster python tools/quick_time.py 11 i8 8192 1.8 > $OUT/ncu_log.txt 2>&1
cat $OUT/log.txt
