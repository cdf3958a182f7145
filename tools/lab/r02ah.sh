#!/bin/bash
# Round-2 lab run AH: TC two-codewords-per-register kernel with the hard decisions inside the messages (MODE 3)
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ah
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_front.py tests/test_gpu_configs.py tests/test_gpu_fuzz.py -x -q -m gpu -k "tc or TC or c1 or i8_awgn or i8_saturation or front or maxiters or fuzz" > $OUT/pytest.txt 2>&1; tail -2 $OUT/pytest.txt
for v in 3 2; do
  for spec in "0 i8 1048576 3.0" "0 i8 262144 0.0" "1 i8 1048576 3.0" "2 i8 524288 2.5"; do
    set -- $spec
    LABRADOR_LDPC_TC_X2_HABS=$v timeout 300 python tools/quick_time.py $1 $2 $3 $4 2>&1 | sed "s/^/x2 mode=$v: /" | cut -c1-180 | tee -a $OUT/log.txt
  done
done
