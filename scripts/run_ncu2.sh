#!/bin/bash
# ncu captures (one GPU): wgrad-shaped 2-CTA GEMM, dGELU-epilogue GEMM
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1"
timeout 280 $NCU -f -o gpurun_out/prof_wgrad_2cta python scripts/gemm_one.py 768 3072 16384 1 0 2 128 > gpurun_out/ncu_wgrad_2cta.log 2>&1; tail -2 gpurun_out/ncu_wgrad_2cta.log
timeout 280 $NCU -f -o gpurun_out/prof_dgelu python scripts/gemm_one.py 16384 3072 768 0 0 1 256 3 > gpurun_out/ncu_dgelu.log 2>&1; tail -2 gpurun_out/ncu_dgelu.log
timeout 280 $NCU -f -o gpurun_out/prof_fwd_2cta python scripts/gemm_one.py 16384 768 3072 0 1 2 256 > gpurun_out/ncu_fwd_2cta.log 2>&1; tail -2 gpurun_out/ncu_fwd_2cta.log
ls -la gpurun_out/*.ncu-rep
