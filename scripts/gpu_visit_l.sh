#!/bin/bash
# r01l visit: GPU parity tests of HEAD, bench (simt gate / tc gate), launch list and full captures of the rows-in-lanes kernels.
set -x
TAG=${1:-r01l}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
HGB_GATE=tc timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_gatetc.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_gatetc.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_tcr -s 8 -c 2 -f -o gpurun_out/${TAG}_tcr_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_tcr.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_tcr.log | cut -c1-200
ls -la gpurun_out/
