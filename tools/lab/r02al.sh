#!/bin/bash
# Round-2 lab run AL: TM8192 i8 with one word slot per thread as the default: whole suite, ncu capture, bench C3.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02al
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -2 $OUT/pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
QUICK_TIME_META=$OUT/r02_tm8192_meta.json timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ms_tm -s 1 -c 1 -f -o $OUT/r02_tm8192 python tools/quick_time.py 8 i8 65536 2.0 > $OUT/ncu_log.txt 2>&1
tail -1 $OUT/ncu_log.txt
timeout 900 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; tail -c 200 $OUT/bench_c3.json; echo
timeout 900 python bench.py --workload c4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err
LABRADOR_LDPC_TM_PROF=1 timeout 300 python tools/quick_time.py 8 i8 16384 2.0 2>&1 | grep -i "tm_prof" | head -2
