#!/bin/bash
# Profiling visit for the default (tcg) message path: launch list + full captures of the small-class message
# kernel, the wide-class message kernel and the radial gate pre-pass (tbg_m8 keeps ncu replays short).
set -x
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msgpack_tcr -s 8 -c 1 -f -o gpurun_out/${TAG}_tcg_small_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_small.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_small.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:radial_gate -s 8 -c 1 -f -o gpurun_out/${TAG}_gate_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_gate.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_gate.log | cut -c1-200
