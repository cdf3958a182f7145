#!/bin/bash
# Round-2 lab run Y: self-correction rule on fp16 lanes (ARITH 8 / 9) against the shipped 632
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02y
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
for a in 832 932; do
  LABRADOR_LDPC_TM_ARITH=$a timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_front.py -q -x -m gpu -k "i8 or c3 or c4 or tm8192 or front" > $OUT/pytest_$a.txt 2>&1; echo "arith $a: $(tail -1 $OUT/pytest_$a.txt)" | tee -a $OUT/log.txt
done
for a in 632 832 932; do
  for spec in "8 i8 65536 2.0" "7 i8 131072 3.0" "6 i8 131072 4.0" "5 i8 262144 2.0" "4 i8 262144 3.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/arith=$a: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
done
