#!/bin/bash
# Round-2 lab run E: TM1280 i8 on the packed-lane kernel (residency variants), sweep over all twelve codes.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02e
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
for v in "LABRADOR_LDPC_TM1280_WIDE=1" "LABRADOR_LDPC_TM1280_MINB=1" "LABRADOR_LDPC_TM1280_MINB=6" "LABRADOR_LDPC_TM1280_MINB=8"; do
  echo "=== $v" >> $OUT/log.txt
  env $v timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_front.py -x -q -k "i8 or maxiters_edge or unaligned or front or soft or hard" >> $OUT/log.txt 2>&1
  env $v timeout 300 python tools/quick_time.py 3 i8 262144 4.0 2>&1 | cut -c1-170 >> $OUT/log.txt
  env $v timeout 300 python tools/quick_time.py 3 i8 262144 3.6 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 1500 python tools/sweep.py > $OUT/sweep_stdout.txt 2>&1
cp gpurun_out/r02_sweep.md $OUT/ 2>/dev/null
grep -v "^\.\.\.\|^$" $OUT/log.txt | tail -30
