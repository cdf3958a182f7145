#!/bin/bash
# Round-2 lab run AF (8 GPUs): C3 and C4 at 2 / 4 / 8 GPUs on the fp16-arithmetic build, the in-library multi-device test.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02af
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k in_process_multi_device_split > $OUT/pytest_multi_device.txt 2>&1; tail -1 $OUT/pytest_multi_device.txt
for n in 8 4 2; do
  run $n --workload c3 --steps 3 > $OUT/bench_c3_${n}gpu.json 2> $OUT/bench_c3_${n}gpu.err; tail -c 200 $OUT/bench_c3_${n}gpu.json; echo
  run $n --workload c4 --steps 3 > $OUT/bench_c4_${n}gpu.json 2> $OUT/bench_c4_${n}gpu.err
done
