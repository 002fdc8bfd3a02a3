#!/bin/bash
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?"; }
T 200 python scripts/gemm_check.py > gpurun_out/gemm_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm_check.log | head -5
T 250 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
T 200 python scripts/profile_step.py ours > gpurun_out/profile_ours.log 2>&1; grep -A12 "^== ours" gpurun_out/profile_ours.log | cut -c1-150
T 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; tail -3 gpurun_out/bench_ours_n1.err; cut -c1-420 gpurun_out/bench_ours_n1.json
