#!/bin/bash
# r01m visit: first GPU run of the edge-aligned ('rot') message path: parity tests, bench, launch list, full capture.
set -x
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_rot.py -x -q -s > gpurun_out/${TAG}_pytest_rot.log 2>&1; RC=$?; tail -25 gpurun_out/${TAG}_pytest_rot.log
if [ $RC -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_rot.py::test_single_message_calls" -x -q -s -k "small and None" > gpurun_out/${TAG}_sanitizer.log 2>&1; tail -40 gpurun_out/${TAG}_sanitizer.log
fi
HGB_MSGPACK=rot timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_rot.json; tail -5 gpurun_out/${TAG}_bench.err
if [ $RC -eq 0 ]; then
HGB_MSGPACK=rot timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
HGB_MSGPACK=rot timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_rot -s 9 -c 3 -f -o gpurun_out/${TAG}_rot_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rot.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_rot.log | cut -c1-200
HGB_MSGPACK=rot timeout 300 ncu --set full --clock-control none --import-source on -k regex:rotate_pack -s 3 -c 1 -f -o gpurun_out/${TAG}_rotpack_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rotpack.log 2>&1
fi
ls -la gpurun_out/
