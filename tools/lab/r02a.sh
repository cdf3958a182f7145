#!/bin/bash
# Round-2 lab run A (one gpurun call): new arithmetic variants of the TM i8 min-sum kernel -- exactness against the
# oracle, device-timed throughput, per-phase cycle profile; resident-CTA sweep on TM2048; ncu of the shipped kernel.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02a
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
for a in 5 6 616 7 716; do
  echo "=== variant $a: parity" >> $OUT/log.txt
  LABRADOR_LDPC_TM_ARITH=$a timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "i8_awgn_exact or i8_saturation_stress or maxiters_edge or full_size_properties" >> $OUT/log.txt 2>&1
  echo "=== variant $a: timing" >> $OUT/log.txt
  for spec in "8 65536 2.0" "7 65536 3.0" "6 65536 4.0" "5 262144 2.0" "4 262144 3.0"; do
    set -- $spec
    LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py $1 i8 $2 $3 2>&1 | sed "s/^/arith $a: /" | cut -c1-160 >> $OUT/log.txt
  done
done
echo "=== phase profiles (TM8192, 16384 frames)" >> $OUT/log.txt
for a in 5 2 4 6 7; do
  LABRADOR_LDPC_TM_PROF=1 LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 8 i8 16384 2.0 2>&1 | grep -E "tm_prof|cw/s" | awk '/tm_prof arith/{n++} n<=1' >> $OUT/log.txt
done
echo "=== resident CTAs per SM, TM2048 / TM1536 (default = as many as fit)" >> $OUT/log.txt
for a in 5 2; do for cap in 1 2 0; do
  LABRADOR_LDPC_TM_CTAS_PER_SM=$cap LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 5 i8 262144 2.0 2>&1 | sed "s/^/arith $a cap $cap: /" | cut -c1-160 >> $OUT/log.txt
  LABRADOR_LDPC_TM_CTAS_PER_SM=$cap LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 4 i8 262144 3.0 2>&1 | sed "s/^/arith $a cap $cap: /" | cut -c1-160 >> $OUT/log.txt
done; done
echo "=== ncu (shipped variant)" >> $OUT/log.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192_a5 python tools/quick_time.py 8 i8 65536 2.0 >> $OUT/log.txt 2>&1
LABRADOR_LDPC_TM_ARITH=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192_a7 python tools/quick_time.py 8 i8 65536 2.0 >> $OUT/log.txt 2>&1
tail -5 $OUT/log.txt
