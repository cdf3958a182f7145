#!/bin/bash
# Round-2 lab run L: ncu of the cluster kernel (TM32768), resident clusters.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02l
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm_cluster -s 1 -c 1 -f -o $OUT/r02_tm32768_cluster python tools/quick_time.py 11 i8 8192 1.8 > $OUT/log.txt 2>&1
tail -3 $OUT/log.txt
