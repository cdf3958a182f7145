#!/bin/bash
# Round-2 lab run O: TC codes, i8: fp16 |v| / minima in the two-codewords-per-register kernel (A/B), whole suite, sweep, C1.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02o
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
for v in 0 1; do
  echo "=== LABRADOR_LDPC_TC_X2_HABS=$v" >> $OUT/log.txt
  LABRADOR_LDPC_TC_X2_HABS=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_front.py tests/test_gpu_configs.py -x -q -k "tc or c1 or i8_awgn or i8_saturation or front or maxiters" >> $OUT/log.txt 2>&1
  for spec in "0 262144 3.0" "1 262144 3.0" "2 262144 2.5" "0 262144 1.0"; do
    set -- $spec
    LABRADOR_LDPC_TC_X2_HABS=$v timeout 300 python tools/quick_time.py $1 i8 $2 $3 2>&1 | cut -c1-170 >> $OUT/log.txt
  done
done
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
timeout 600 python bench.py --workload c1 > $OUT/bench_c1.json 2> $OUT/bench_c1.err
timeout 1500 python tools/sweep.py > $OUT/sweep_stdout.txt 2>&1; cp gpurun_out/r02_sweep.md $OUT/ 2>/dev/null
grep -v "^\.\.\.\|^$" $OUT/log.txt
