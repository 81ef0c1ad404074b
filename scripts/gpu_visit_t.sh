#!/bin/bash
set -x
TAG=${1:-r01u}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_rot.py -q -s -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; RC=$?; grep -E "rot gate|rot chunk|passed|failed|Error" gpurun_out/${TAG}_pytest_gpu.log | tail -12
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot.json 2> gpurun_out/${TAG}_bench.err; cut -c1-330 gpurun_out/${TAG}_bench_rot.json; tail -5 gpurun_out/${TAG}_bench.err
HGB_GATE=tc timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot_gatetc.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-330 gpurun_out/${TAG}_bench_rot_gatetc.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_rot -s 9 -c 2 -f -o gpurun_out/${TAG}_rot_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rot.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_rot.log | cut -c1-200
ls -la gpurun_out/ | tail -6
