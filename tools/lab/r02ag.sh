#!/bin/bash
# Round-2 lab run AG: compute-sanitizer (memcheck, racecheck) over every kernel family of the final build; perftest and the
# reference-bench harness still run.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ag
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $OUT/log.txt; tail -2 $OUT/memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > $OUT/racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $OUT/log.txt; tail -2 $OUT/racecheck.txt
timeout 600 python tools/bench_decode_rs.py > $OUT/bench_decode_rs.txt 2>&1; tail -12 $OUT/bench_decode_rs.txt
timeout 600 python perftest.py --help > /dev/null 2>&1; echo "perftest help rc=$?"
