#!/bin/bash
# Round-2 lab run AM: ARITH 11 (lane swap + fp16 exponent by one PRMT) against ARITH 10
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02am
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
LABRADOR_LDPC_TM_ARITH=1132 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -m gpu -k "i8 or c3 or c4 or tm8192" > $OUT/pytest_1132.txt 2>&1; echo "arith 1132: $(tail -1 $OUT/pytest_1132.txt)" | tee -a $OUT/log.txt
for a in 1032 1132; do
  for spec in "8 i8 65536 2.0" "7 i8 131072 3.0" "5 i8 262144 2.0" "4 i8 262144 3.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/arith=$a: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
  LABRADOR_LDPC_TM_WPT=2 LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 8 i8 65536 2.0 2>&1 | sed "s/^/arith=$a wpt=2: /" | cut -c1-180 | tee -a $OUT/log.txt
done
for a in 10322 11322; do LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 6 i8 131072 4.0 2>&1 | sed "s/^/arith=$a: /" | cut -c1-180 | tee -a $OUT/log.txt; done
