#!/bin/bash
# cluster kernel: st.async + mbarrier (no cluster barrier inside an iteration) against the two-barrier version
export LABRADOR_LDPC_NO_REBUILD=1
mkdir -p gpurun_out/r02t
timeout 600 python -m pytest tests/test_gpu_k16384.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02t/pytest_async.txt
cat gpurun_out/r02t/pytest_async.txt
for v in 1 0; do for c in "11 1.8" "10 2.6" "9 3.6"; do set -- $c
  LABRADOR_LDPC_CLUSTER_ASYNC=$v timeout 200 python tools/quick_time.py $1 i8 8192 $2 2>&1 | sed "s/^/async=$v: /" | cut -c1-175
done; done | tee gpurun_out/r02t/log.txt
LABRADOR_LDPC_CLUSTER_LOCAL_PUSH=1 timeout 200 python tools/quick_time.py 11 i8 8192 1.8 2>&1 | sed "s/^/async=1 local_push=1: /" | cut -c1-175 | tee -a gpurun_out/r02t/log.txt
