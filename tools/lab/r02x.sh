#!/bin/bash
# Round-2 lab run X: checkpoint of the whole tree -- GPU suite, smoke(), sweep of every (operation, code, type), C3 bench line.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02x
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
timeout 1200 python tools/sweep.py > $OUT/sweep.md 2> $OUT/sweep.err; tail -14 $OUT/sweep.md | cut -c1-200
timeout 600 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; tail -c 400 $OUT/bench_c3.json
