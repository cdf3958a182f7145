#!/bin/bash
# Round-2 lab run AE: every bench workload at N=1 on the fp16-arithmetic build, the reference arm, the sweep, smoke().
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ae
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
for w in c3 c1 c2 c4 c5; do
  timeout 900 python bench.py --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -c 300 $OUT/bench_$w.json; echo
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_c3.json 2>&1
timeout 1200 python tools/sweep.py > $OUT/sweep.md 2> $OUT/sweep.err; tail -13 $OUT/sweep.md | cut -c1-150
