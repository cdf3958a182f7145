#!/bin/bash
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ad
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 900 python tools/lab/cluster_stress.py 2 2>&1 | tail -12 | tee $OUT/stress.txt
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for spec in "8 i8 65536 2.0" "5 i8 262144 2.0" "0 i8 1048576 3.0" "11 i8 8192 1.8"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 | tee -a $OUT/log.txt
done
QUICK_TIME_META=$OUT/r02_tm8192_meta.json timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192 python tools/quick_time.py 8 i8 65536 2.0 > $OUT/ncu_log.txt 2>&1
tail -2 $OUT/ncu_log.txt
