#!/bin/bash
# Round-2 lab run W: quad + st.async cluster kernel -- variant tests and one ncu capture (TM32768)
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02w
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 900 python -m pytest tests/test_gpu_k16384.py -q -x > $OUT/pytest_k16384.txt 2>&1; echo "k16384 rc=$?" >> $OUT/log.txt; tail -3 $OUT/pytest_k16384.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm_cluster4 -s 1 -c 1 -f -o $OUT/r02_tm32768_cluster4 python tools/quick_time.py 11 i8 8192 1.8 > $OUT/ncu_log.txt 2>&1
tail -2 $OUT/ncu_log.txt
cat $OUT/log.txt
