#!/bin/bash
# rot-path visit: parity tests, bench variants, launch list, full captures.
set -x
TAG=${1:-r01n}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_rot.py -q -s > gpurun_out/${TAG}_pytest_rot.log 2>&1; RC=$?; grep -E "rel err|passed|failed" gpurun_out/${TAG}_pytest_rot.log | tail -25
HGB_MSGPACK=rot timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_rot.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_rot.json; tail -5 gpurun_out/${TAG}_bench.err
HGB_MSGPACK=rot HGB_GATE=tc timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot_gatetc.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-330 gpurun_out/${TAG}_bench_rot_gatetc.json
HGB_MSGPACK=rot HGB_ROT_CHUNK=32768 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot_c32k.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-330 gpurun_out/${TAG}_bench_rot_c32k.json
HGB_MSGPACK=rot HGB_ROT_CHUNK=524288 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rot_c512k.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-330 gpurun_out/${TAG}_bench_rot_c512k.json
if [ $RC -eq 0 ]; then
HGB_MSGPACK=rot timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
HGB_MSGPACK=rot timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_rot -s 9 -c 3 -f -o gpurun_out/${TAG}_rot_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rot.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_rot.log | cut -c1-200
HGB_MSGPACK=rot timeout 300 ncu --set full --clock-control none --import-source on -k regex:rotate_pack -s 3 -c 1 -f -o gpurun_out/${TAG}_rotpack_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_rotpack.log 2>&1
fi
ls -la gpurun_out/
