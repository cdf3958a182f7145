#!/bin/bash
# Round-2 lab run J: scalar-lane kernel with TMA-staged frames (tests + timing); reference benches again (CPU side fixed).
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02j
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for spec in "5 f32 262144 2.0" "5 i32 262144 2.0" "5 f64 65536 2.0" "8 f32 32768 2.0" "3 f32 262144 4.0" "7 f32 65536 3.0" "6 f64 32768 4.0" "3 i16 262144 4.0"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 900 python tools/bench_decode_rs.py > $OUT/benches_stdout.txt 2>&1; cp gpurun_out/r02_reference_benches.md $OUT/ 2>/dev/null
cat $OUT/log.txt; tail -10 $OUT/benches_stdout.txt
