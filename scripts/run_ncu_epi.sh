#!/bin/bash
# ncu captures (one GPU) of the two epilogue-heavy GEMMs of the GPT-2 step on the ring-epilogue pair kernel
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1"
timeout 280 $NCU -f -o gpurun_out/prof_gelu_ring python scripts/gemm_one.py 16384 3072 768 0 0 0 0 1 > gpurun_out/ncu_gelu_ring.log 2>&1; tail -2 gpurun_out/ncu_gelu_ring.log
timeout 280 $NCU -f -o gpurun_out/prof_dgelu_ring python scripts/gemm_one.py 16384 3072 768 0 1 0 0 3 > gpurun_out/ncu_dgelu_ring.log 2>&1; tail -2 gpurun_out/ncu_dgelu_ring.log
