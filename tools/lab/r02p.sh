#!/bin/bash
# Round-2 lab run P: scalar-lane kernel with three-input minima (f32 / i32), smoke(), whole suite, C2 / C3 bench lines.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02p
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
for spec in "5 f32 262144 2.0" "5 i32 262144 2.0" "8 f32 32768 2.0" "3 f32 262144 4.0" "7 f32 65536 3.0" "6 f32 65536 4.0" "6 i32 65536 4.0"; do
  set -- $spec
  timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | cut -c1-170 >> $OUT/log.txt
done
timeout 600 python bench.py --workload c2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
timeout 600 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_c3_ref.json 2>&1
cat $OUT/log.txt; tail -c 300 $OUT/bench_c3.json
