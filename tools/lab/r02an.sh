#!/bin/bash
# Round-2 lab run AN: last check of the tree as committed -- whole GPU suite, smoke(), default bench line.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02an
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -2 $OUT/pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
timeout 900 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; python - <<PY
import json
j=json.loads(open("$OUT/bench_c3.json").read().strip().splitlines()[-1])
print(j["value"], j["unit"], "e2e", j["e2e"]["value"], "roofline", j["roofline"]["bound"], round(j["roofline"]["frac"],3), "parity", j["parity"]["frames_compared"], j["parity"]["mismatches"], "cpu", j["cpu_baseline"]["value"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 300
