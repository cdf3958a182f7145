#!/bin/bash
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ac
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 900 python tools/lab/cluster_stress.py 4 2>&1 | tail -30 | tee $OUT/stress.txt
timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest.txt 2>&1; tail -6 $OUT/pytest.txt
