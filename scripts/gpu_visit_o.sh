#!/bin/bash
set -x
TAG=${1:-r01o}
mkdir -p gpurun_out
timeout 400 python scripts/error_budget.py > gpurun_out/${TAG}_error_budget.log 2>&1; cat gpurun_out/${TAG}_error_budget.log | tail -20
HGB_MSGPACK=rot timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
HGB_MSGPACK=rot timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_rot -s 9 -c 3 -f -o gpurun_out/${TAG}_rot_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rot.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_rot.log | cut -c1-200
HGB_MSGPACK=rot timeout 300 ncu --set full --clock-control none --import-source on -k regex:rotate_pack -s 3 -c 1 -f -o gpurun_out/${TAG}_rotpack_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rotpack.log 2>&1
ls -la gpurun_out/
