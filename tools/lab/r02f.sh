#!/bin/bash
# Round-2 lab run F: i16 on the packed-check-side kernel (decode_ms_tm_i16.cu) against the scalar-lane kernel.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02f
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_front.py tests/test_gpu_configs.py -x -q -k "i16 or wide_int or front or soft or c2" > $OUT/pytest_i16.txt 2>&1; tail -8 $OUT/pytest_i16.txt
for v in "LABRADOR_LDPC_TM_I16_WIDE=1" "LABRADOR_LDPC_TM_I16_WIDE=0"; do
  echo "=== $v" >> $OUT/log.txt
  for spec in "5 262144 2.0" "4 262144 3.0" "6 65536 4.0" "7 65536 3.0" "8 65536 2.0"; do
    set -- $spec
    env $v timeout 300 python tools/quick_time.py $1 i16 $2 $3 2>&1 | cut -c1-170 >> $OUT/log.txt
  done
done
cat $OUT/log.txt
