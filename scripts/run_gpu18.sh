#!/bin/bash
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?" >&2; }
T 100 python scripts/attn_layout_probe.py 2>&1 | tail -8
T 200 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; grep -E "ALL_OK|FAIL|Error" gpurun_out/fused_check.log | head -5
T 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
T 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | sed -n 2,3p
T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; tail -2 gpurun_out/bench_ours_n1.err; cut -c1-330 gpurun_out/bench_ours_n1.json
