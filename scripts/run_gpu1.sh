#!/bin/bash
# single-GPU validation + bench (gpurun --timeout 900 -- 'bash scripts/run_gpu1.sh'):
# kernel numerics, GPU unit tests, per-kernel trace of one step, both bench arms.
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?" >&2; }
T 200 python scripts/gemm_check.py > gpurun_out/gemm_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm_check.log | head -5
TDP_GEMM_EPI=split TDP_GEMM_2CTA=0 T 200 python scripts/gemm_check.py > gpurun_out/gemm_check_epi_split.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm_check_epi_split.log | head -5
T 150 python scripts/gemm2cta_check.py > gpurun_out/gemm2cta_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm2cta_check.log | tail -3
T 150 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; grep -E "all_ok|FAIL|Error" gpurun_out/fused_check.log | head -3
T 150 python scripts/grouped_check.py > gpurun_out/grouped_check.log 2>&1; grep -E "ALL_OK|'ok': False|error" gpurun_out/grouped_check.log | tail -4 | cut -c1-300
T 150 python scripts/attn_check.py > gpurun_out/attn_check.log 2>&1; grep -E "ALL_OK|'ok': False|ERROR|native_ms" gpurun_out/attn_check.log | tail -6
T 250 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
T 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | sed -n 2,3p
T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; cut -c1-330 gpurun_out/bench_ours_n1.json
TDP_GEMM_EPI=split T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1_epi_split.json 2> gpurun_out/bench_ours_n1_epi_split.err; cut -c1-330 gpurun_out/bench_ours_n1_epi_split.json
TDP_ATTN=native T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1_attn_native.json 2> gpurun_out/bench_ours_n1_attn_native.err; cut -c1-330 gpurun_out/bench_ours_n1_attn_native.json
T 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; cut -c1-330 gpurun_out/bench_ref_n1.json
