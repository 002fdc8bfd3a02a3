#!/bin/bash
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?"; }
T 200 python scripts/wgrad_bench.py 2>&1 | tail -8
T 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 -s 7 -c 3 -f -o gpurun_out/prof_wgrad python scripts/wgrad_bench.py --one > gpurun_out/ncu_wgrad.log 2>&1; tail -3 gpurun_out/ncu_wgrad.log
