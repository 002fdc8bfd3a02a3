#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?" >&2; }
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?" >&2; }
T 150 python scripts/gemm2cta_check.py > gpurun_out/gemm2cta_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error|cublas_ms|dgelu" gpurun_out/gemm2cta_check.log | cut -c1-330 | tail -28
T 200 python scripts/gemm_check.py > gpurun_out/gemm_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error" gpurun_out/gemm_check.log | head -5
T 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
T 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | sed -n 2,3p
T 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; tail -2 gpurun_out/bench_ours_n1.err; cut -c1-330 gpurun_out/bench_ours_n1.json
if [ "$N" -gt 1 ]; then
TDP_SYMM_BACKEND=native TR 200 29511 scripts/symm_check.py > gpurun_out/symm_check_native_w$N.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/symm_check_native_w$N.log | grep -E "multicast|ALL_OK|FAIL|Error|gemm_rs|ag_gemm|'MiB': 25" | cut -c1-420 | tail -8
TR 200 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/tp_check_w$N.log | tail -8 | cut -c1-330
for impl in ours ours_nccl reference; do
  TR 150 29520 scripts/bench_tp.py --impl $impl > gpurun_out/bench_tp_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_tp_${impl}_w$N.log | tail -2 | cut -c1-330
done
fi
