#!/bin/bash
export LABRADOR_LDPC_NO_REBUILD=1
mkdir -p gpurun_out/r02v
LABRADOR_LDPC_CLUSTER_ASYNC=1 timeout 600 python -m pytest tests/test_gpu_k16384.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02v/pytest_quad_async.txt
cat gpurun_out/r02v/pytest_quad_async.txt
for v in 1 0; do
 for c in "11 1.8" "10 2.6" "9 3.6"; do set -- $c
  LABRADOR_LDPC_CLUSTER_ASYNC=$v timeout 200 python tools/quick_time.py $1 i8 8192 $2 2>&1 | sed "s/^/quad async=$v: /" | cut -c1-185
done; done | tee gpurun_out/r02v/log.txt
