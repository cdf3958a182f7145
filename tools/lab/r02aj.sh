#!/bin/bash
# Round-2 lab run AJ: final state -- whole GPU suite, smoke(), bench C1 / C3, sweep.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02aj
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
for w in c1 c3; do
  timeout 900 python bench.py --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; python - <<PY
import json
j=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
print("$w", j["value"], j["unit"], "e2e", j["e2e"]["value"], "parity", j["parity"]["frames_compared"], j["parity"]["mismatches"], "launches", j["gpu_launches"])
PY
done
timeout 1200 python tools/sweep.py > $OUT/sweep.md 2> $OUT/sweep.err; tail -13 $OUT/sweep.md | cut -c1-120
