#!/bin/bash
# ring-epilogue GEMM variant: numerics + step time with / without it (one GPU)
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?" >&2; }
T 150 python scripts/gemm2cta_check.py > gpurun_out/gemm2cta_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm2cta_check.log | tail -3
T 200 python scripts/gemm_check.py > gpurun_out/gemm_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm_check.log | head -5
T 150 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; grep -E "all_ok|BAD" gpurun_out/fused_check.log | head -3
T 300 python -m pytest tests -m gpu -x -q -k "gemm or mlp or gpt2 or linear" 2>&1 | tail -3
T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; cut -c1-200 gpurun_out/bench_ours_n1.json
TDP_GEMM_EPIRING=0 T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1_noring.json 2> gpurun_out/bench_ours_n1_noring.err; cut -c1-200 gpurun_out/bench_ours_n1_noring.json
T 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | sed -n 2,3p
