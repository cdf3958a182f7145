#!/bin/bash
# Round-2 lab run C: full GPU test suite on the new default variant + runtime, every bench workload at N=1,
# ncu capture (with launch facts) of the shipped headline kernel, two more variant timings.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02c
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for w in c3 c1 c2 c4 c5; do
  timeout 600 python bench.py --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -c 600 $OUT/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_c3.json 2>&1
for a in 6322 632; do LABRADOR_LDPC_TM_ARITH=$a timeout 300 python tools/quick_time.py 6 i8 65536 4.0 2>&1 | sed "s/^/arith $a: /" | cut -c1-160 >> $OUT/log.txt; done
LABRADOR_LDPC_TM_WPT=1 timeout 300 python tools/quick_time.py 8 i8 65536 2.0 2>&1 | sed "s/^/wpt1 (632): /" | cut -c1-160 >> $OUT/log.txt
QUICK_TIME_META=$OUT/r02_tm8192_meta.json timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192 python tools/quick_time.py 8 i8 65536 2.0 >> $OUT/log.txt 2>&1
python tools/latency.py > $OUT/latency.txt 2>&1
tail -5 $OUT/log.txt
gcc -O2 -pthread -Iinclude tools/calls_per_s.c -o /tmp/calls_per_s -Llabrador_ldpc_b200/lib -llabrador_ldpc -Wl,-rpath,$PWD/labrador_ldpc_b200/lib 2>> $OUT/log.txt
for code in 8 0; do for t in 1 2 4 8 16 32; do /tmp/calls_per_s $code $t 3000 >> $OUT/calls.txt 2>&1; done; done
cat $OUT/calls.txt
