#!/bin/bash
# Round-2 lab run G (8 GPUs): 1 -> 8 curves of the C4 and C5 workloads, C3 at 8 (with the copy-only control and the
# library's own multi-device split), the in-library multi-device test.
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02g
mkdir -p $OUT
export LABRADOR_LDPC_NO_REBUILD=1
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > $OUT/cpu.txt
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k in_process_multi_device_split > $OUT/pytest_multi_device.txt 2>&1; tail -2 $OUT/pytest_multi_device.txt
run 8 --workload c3 --steps 3 > $OUT/bench_c3_8gpu.json 2> $OUT/bench_c3_8gpu.err; tail -c 300 $OUT/bench_c3_8gpu.json
for n in 2 4 8; do
  run $n --workload c4 --steps 3 > $OUT/bench_c4_${n}gpu.json 2> $OUT/bench_c4_${n}gpu.err
  run $n --workload c5 --steps 3 > $OUT/bench_c5_${n}gpu.json 2> $OUT/bench_c5_${n}gpu.err
done
run 2 --workload c3 --steps 3 > $OUT/bench_c3_2gpu.json 2> $OUT/bench_c3_2gpu.err
ls -la $OUT
