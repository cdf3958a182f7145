#!/bin/bash
# cluster kernel: four messages per word against two
export LABRADOR_LDPC_NO_REBUILD=1
mkdir -p gpurun_out/r02u
timeout 600 python -m pytest tests/test_gpu_k16384.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02u/pytest_quad.txt
cat gpurun_out/r02u/pytest_quad.txt
LABRADOR_LDPC_CLUSTER_MINB=1 timeout 600 python -m pytest tests/test_gpu_k16384.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02u/pytest_quad_minb1.txt
cat gpurun_out/r02u/pytest_quad_minb1.txt
for v in "1 2" "1 1" "0 2"; do set -- $v; qd=$1; mb=$2
 for c in "11 1.8" "10 2.6" "9 3.6"; do set -- $c
  LABRADOR_LDPC_CLUSTER_QUAD=$qd LABRADOR_LDPC_CLUSTER_MINB=$mb timeout 200 python tools/quick_time.py $1 i8 8192 $2 2>&1 | sed "s/^/quad=$qd minb=$mb: /" | cut -c1-185
done; done | tee gpurun_out/r02u/log.txt
