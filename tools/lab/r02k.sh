#!/bin/bash
# Round-2 lab run K: the cluster kernel of the k = 16384 codes (tests under a tight timeout first), TM1280 i16 packed.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02k
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 240 python -m pytest tests/test_gpu_k16384.py -q -x -k "clean or encoder" > $OUT/pytest_first.txt 2>&1; echo "first rc=$?" >> $OUT/log.txt; tail -15 $OUT/pytest_first.txt
nvidia-smi --query-gpu=name,memory.used --format=csv >> $OUT/log.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_k16384.py -q -x > $OUT/pytest_k16384.txt 2>&1; echo "k16384 rc=$?" >> $OUT/log.txt; tail -15 $OUT/pytest_k16384.txt
for spec in "11 i8 8192 1.8" "10 i8 8192 2.6" "9 i8 8192 3.6" "11 i8 65536 1.8"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_front.py -q -x -k "i16 or dispatch or front" > $OUT/pytest_i16.txt 2>&1; tail -3 $OUT/pytest_i16.txt
timeout 300 python tools/quick_time.py 3 i16 262144 4.0 2>&1 | cut -c1-170 >> $OUT/log.txt
cat $OUT/log.txt
