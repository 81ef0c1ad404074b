#!/bin/bash
# Quick GPU visit for the tensor-core path: building-block self-test, parity of the tcgen05 message kernel,
# small and full benches for both message kernels.  Tight per-stage timeouts.
set -x
TAG=${1:-r01b}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_tcgen05.py -x -q -s > gpurun_out/${TAG}_tc_selftest.log 2>&1; tail -15 gpurun_out/${TAG}_tc_selftest.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k tensor_core > gpurun_out/${TAG}_tc_parity.log 2>&1; tail -15 gpurun_out/${TAG}_tc_parity.log
HGB_MSGPACK=simt timeout 300 python bench.py --steps 3 --warmup 3 --workload tbg_m8 --no-cpu-baseline > gpurun_out/${TAG}_bench_m8_simt.json 2> gpurun_out/${TAG}_bench_m8_simt.err; cut -c1-700 gpurun_out/${TAG}_bench_m8_simt.json; tail -3 gpurun_out/${TAG}_bench_m8_simt.err
HGB_MSGPACK=tc timeout 300 python bench.py --steps 3 --warmup 3 --workload tbg_m8 --no-cpu-baseline > gpurun_out/${TAG}_bench_m8_tc.json 2> gpurun_out/${TAG}_bench_m8_tc.err; cut -c1-700 gpurun_out/${TAG}_bench_m8_tc.json; tail -3 gpurun_out/${TAG}_bench_m8_tc.err
HGB_MSGPACK=tc timeout 400 python bench.py --steps 3 --warmup 3 --workload tbg_m28 > gpurun_out/${TAG}_bench_m28_tc.json 2> gpurun_out/${TAG}_bench_m28_tc.err; cut -c1-1500 gpurun_out/${TAG}_bench_m28_tc.json; tail -3 gpurun_out/${TAG}_bench_m28_tc.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
