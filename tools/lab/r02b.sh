#!/bin/bash
# Round-2 lab run B: in-thread row-0 exit test (KNOBS bit 5) on top of ARITH 5 / 6.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02b
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
for a in 632 532 6; do
  echo "=== variant $a: parity" >> $OUT/log.txt
  LABRADOR_LDPC_TM_ARITH=$a timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_front.py -x -q -k "i8 or maxiters_edge or full_size_properties or c3 or c4 or unaligned or mixed" >> $OUT/log.txt 2>&1
  echo "=== variant $a: timing" >> $OUT/log.txt
  for spec in "8 65536 2.0" "7 65536 3.0" "6 65536 4.0" "5 262144 2.0" "4 262144 3.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 i8 $2 $3 2>&1 | sed "s/^/arith $a: /" | cut -c1-160 >> $OUT/log.txt
  done
done
echo "=== phase profile 632" >> $OUT/log.txt
LABRADOR_LDPC_TM_PROF=1 LABRADOR_LDPC_TM_ARITH=632 timeout 300 python tools/quick_time.py 8 i8 16384 2.0 2>&1 | grep -E "tm_prof|cw/s" | awk '/tm_prof arith/{n++} n<=1' >> $OUT/log.txt
echo "=== ncu 632" >> $OUT/log.txt
LABRADOR_LDPC_TM_ARITH=632 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192_a632 python tools/quick_time.py 8 i8 65536 2.0 >> $OUT/log.txt 2>&1
tail -3 $OUT/log.txt
