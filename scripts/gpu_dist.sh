#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): correctness of the edge-sharded forward vs the unsharded one, then bench.py
# at N ranks launched exactly like the driver does.
set -x
N=${1:-2}
TAG=${2:-r01h}
BACKEND=${3:-rot}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
HGB_MSGPACK=$BACKEND timeout 240 $TR scripts/dist_gpu_check.py > gpurun_out/${TAG}_dist_check_n$N.log 2>&1; grep -E "rank|DIST_CHECK|Error|error" gpurun_out/${TAG}_dist_check_n$N.log | cut -c1-300 | tail -8
HGB_MSGPACK=$BACKEND timeout 400 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload tbg_m28 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; cut -c1-1500 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err | cut -c1-300
