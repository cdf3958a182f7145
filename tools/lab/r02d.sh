#!/bin/bash
# Round-2 lab run D: the k = 16384 codes (tests + timing), whole GPU suite, host-pipeline chunk sizes for the e2e leg.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02d
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 1200 python -m pytest tests/test_gpu_k16384.py -q -x > $OUT/pytest_k16384.txt 2>&1; tail -15 $OUT/pytest_k16384.txt
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_k16384.py > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for mb in 16 32 64 128; do
  LABRADOR_LDPC_CHUNK_MB=$mb timeout 600 python bench.py --workload c3 --frames-per-gpu 262144 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('chunk_mb $mb: value %.2f e2e %.2f copy-control h2d %.1f GB/s (transport-only %.1f Gbit/s)' % (d['value'], e['value'], e['copy_control']['h2d_gbs_per_gpu'], e['copy_control']['value_if_transport_only']))" >> $OUT/log.txt
done
cat $OUT/log.txt
