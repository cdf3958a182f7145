#!/bin/bash
# Round-2 lab run H: whole GPU suite after the scalar-lane kernel got the in-thread exit test; C2 bench; f32 / i32 / f64 timing.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02h
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for spec in "5 f32 262144 2.0" "5 i32 262144 2.0" "5 f64 65536 2.0" "8 f32 32768 2.0" "3 f32 262144 4.0" "7 f32 65536 3.0" "6 f64 32768 4.0"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 600 python bench.py --workload c2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
timeout 600 python bench.py --workload c5 --steps 20 > $OUT/bench_c5.json 2> $OUT/bench_c5.err
cat $OUT/log.txt; tail -c 400 $OUT/bench_c2.json
