#!/bin/bash
# Round-2 lab run I: whole GPU suite; calls per second with lane-owned work counters; the reference's benches as a
# device-timed harness; compute-sanitizer over every kernel family incl. the new ones.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02i
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
gcc -O2 -pthread -Iinclude tools/calls_per_s.c -o /tmp/calls_per_s -Llabrador_ldpc_b200/lib -llabrador_ldpc -Wl,-rpath,$PWD/labrador_ldpc_b200/lib 2>> $OUT/log.txt
for code in 8 0; do for t in 1 2 4 8 16 32; do /tmp/calls_per_s $code $t 3000 >> $OUT/calls.txt 2>&1; done; done
cat $OUT/calls.txt
timeout 900 python tools/bench_decode_rs.py > $OUT/benches_stdout.txt 2>&1; cp gpurun_out/r02_reference_benches.md $OUT/ 2>/dev/null; tail -12 $OUT/benches_stdout.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $OUT/log.txt; tail -3 $OUT/memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > $OUT/racecheck.txt 2>&1; echo "racecheck rc=$?" >> $OUT/log.txt; tail -3 $OUT/racecheck.txt
cat $OUT/log.txt
