#!/bin/bash
# Round-2 lab run Z: ARITH 9 (fp16 self-correction) as the default of every TM code: i8 / front tests, timings, ncu capture.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02z
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_front.py tests/test_gpu_fuzz.py -q -x -m gpu > $OUT/pytest.txt 2>&1; tail -2 $OUT/pytest.txt
for spec in "8 i8 65536 2.0" "7 i8 131072 3.0" "6 i8 131072 4.0" "5 i8 262144 2.0" "4 i8 262144 3.0" "3 i8 524288 4.0"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 | tee -a $OUT/log.txt
done
LABRADOR_LDPC_TM_ARITH=632 timeout 300 python tools/quick_time.py 3 i8 524288 4.0 2>&1 | sed "s/^/arith=632: /" | cut -c1-180 | tee -a $OUT/log.txt
LABRADOR_LDPC_TM_ARITH=6322 timeout 300 python tools/quick_time.py 6 i8 131072 4.0 2>&1 | sed "s/^/arith=6322: /" | cut -c1-180 | tee -a $OUT/log.txt
LABRADOR_LDPC_TM_ARITH=932 timeout 300 python tools/quick_time.py 6 i8 131072 4.0 2>&1 | sed "s/^/arith=932: /" | cut -c1-180 | tee -a $OUT/log.txt
LABRADOR_LDPC_TM_PROF=1 timeout 300 python tools/quick_time.py 8 i8 16384 2.0 2>&1 | grep -i "tm_prof" | head -8 | tee -a $OUT/log.txt
QUICK_TIME_META=$OUT/r02_tm8192_meta.json timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192 python tools/quick_time.py 8 i8 65536 2.0 > $OUT/ncu_log.txt 2>&1
tail -2 $OUT/ncu_log.txt
