#!/bin/bash
# Round-2 lab run AK: ncu launch list of the bench command (shares of the step), short perftest run.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ak
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --frames-per-gpu 262144 --e2e-bytes 268435456 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
tail -c 300 $OUT/bench_under_ncu.log; echo
wc -l $OUT/launches.csv
timeout 600 python perftest.py --help 2>&1 | head -20 > $OUT/perftest_help.txt
timeout 600 python perftest.py --code TM8192 --llr i8 --snrs 1.6,1.8,2.0 --max-bits 2e8 > $OUT/perftest_tm8192_i8.csv 2> $OUT/perftest.err; cat $OUT/perftest_tm8192_i8.csv | head -8
