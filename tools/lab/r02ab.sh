#!/bin/bash
# Round-2 lab run AB: ARITH 10 default on the TM codes, the same arithmetic in the TC two-codewords-per-register kernel
# (MODE 2) and in the cluster kernel: whole GPU suite, timings, ncu capture of the headline kernel.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ab
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1800 python -m pytest tests -q -x -m gpu > $OUT/pytest.txt 2>&1; tail -2 $OUT/pytest.txt
for v in 2 1; do
  for spec in "0 i8 1048576 3.0" "0 i8 262144 0.0" "1 i8 1048576 3.0" "2 i8 524288 2.5"; do
    set -- $spec
    LABRADOR_LDPC_TC_X2_HABS=$v timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/x2 mode=$v: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
done
for spec in "11 i8 8192 1.8" "10 i8 8192 2.6" "9 i8 8192 3.6" "8 i8 65536 2.0" "6 i8 131072 4.0" "3 i8 524288 4.0"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 | tee -a $OUT/log.txt
done
LABRADOR_LDPC_TM_PROF=1 timeout 300 python tools/quick_time.py 8 i8 16384 2.0 2>&1 | grep -i "tm_prof" | head -3 | tee -a $OUT/log.txt
QUICK_TIME_META=$OUT/r02_tm8192_meta.json timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192 python tools/quick_time.py 8 i8 65536 2.0 > $OUT/ncu_log.txt 2>&1
tail -2 $OUT/ncu_log.txt
