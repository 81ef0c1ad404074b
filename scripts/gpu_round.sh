#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list, ncu full capture of the fused
# message kernel.  Everything lands in gpurun_out/ (merged back by gpurun).
set -x
TAG=${1:-r01}
WL=${2:-tbg_m28}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --workload $WL > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --workload $WL > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msgpack -s 8 -c 2 -f -o gpurun_out/${TAG}_msgpack_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
