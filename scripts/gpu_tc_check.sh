#!/bin/bash
# GPU visit: parity prints, tensor-core / big-tile SIMT message kernel benches + ncu.  Tight per-stage timeouts.
set -x
TAG=${1:-r01j}
BK=${2:-tcg}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -s -m gpu -k "tensor_core or full_forward or golden" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
grep -E "rel err|passed|failed" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
if grep -q "passed" gpurun_out/${TAG}_pytest_gpu.log; then
  HGB_MSGPACK=$BK timeout 120 python bench.py --steps 3 --warmup 3 --workload tbg_m8 --no-cpu-baseline > gpurun_out/${TAG}_bench_m8_$BK.json 2> gpurun_out/${TAG}_bench_m8_$BK.err; cut -c1-200 gpurun_out/${TAG}_bench_m8_$BK.json; tail -3 gpurun_out/${TAG}_bench_m8_$BK.err
  HGB_MSGPACK=$BK timeout 300 python bench.py --steps 3 --warmup 3 --workload tbg_m28 --no-cpu-baseline > gpurun_out/${TAG}_bench_m28_$BK.json 2> gpurun_out/${TAG}_bench_m28_$BK.err; cut -c1-1800 gpurun_out/${TAG}_bench_m28_$BK.json; tail -3 gpurun_out/${TAG}_bench_m28_$BK.err
  HGB_MSGPACK=$BK timeout 400 ncu --set full --clock-control none --import-source on -k regex:msgpack_tcr -s 8 -c 1 -f -o gpurun_out/${TAG}_msgpack_${BK}_full \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_tc.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_tc.log | cut -c1-300
fi
if grep -q "passed" gpurun_out/${TAG}_pytest_gpu.log; then
  HGB_MSGPACK=$BK timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_m8.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload tbg_m8 > gpurun_out/${TAG}_ncu_list.log 2>&1
fi
