#!/bin/bash
# Round-2 lab run AO (8 GPUs): C3 at eight GPUs on the final build (one word slot per thread).
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02ao
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload c3 --steps 3 > $OUT/bench_c3_8gpu.json 2> $OUT/bench_c3_8gpu.err; tail -c 200 $OUT/bench_c3_8gpu.json
