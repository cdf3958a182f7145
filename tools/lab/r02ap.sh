#!/bin/bash
# Round-2 lab run AP: iteration 0 with its own copy of both phases (KNOBS bit 6) against the shipped kernel
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ap
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
LABRADOR_LDPC_TM_ARITH=1096 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -m gpu -k "i8 or c3 or c4 or tm8192" > $OUT/pytest_1096.txt 2>&1; echo "arith 1096: $(tail -1 $OUT/pytest_1096.txt)" | tee -a $OUT/log.txt
for a in 1032 1096 1032 1096; do
  for spec in "8 i8 65536 2.0" "5 i8 262144 2.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/arith=$a: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
done
for a in 1032 1096; do
  for spec in "7 i8 131072 3.0" "4 i8 262144 3.0" "3 i8 524288 4.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/arith=$a: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
done
