#!/bin/bash
# ncu --set full of the bandwidth-bound kernels (one GPU)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'layernorm|cross_entropy|adamw_kernel|colsum|rows_copy|sumsq_multi|scale_multi' -c 12 -f -o gpurun_out/prof_misc python scripts/misc_kernels_one.py > gpurun_out/ncu_misc.log 2>&1; tail -2 gpurun_out/ncu_misc.log
timeout 200 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; grep -E "all_ok|BAD" gpurun_out/fused_check.log | head -3
