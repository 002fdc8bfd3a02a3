#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?" >&2; }
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?" >&2; }
T 150 python scripts/gemm2cta_check.py > gpurun_out/gemm2cta_check.log 2>&1; grep -E "ALL_OK|'ok': False|Error|wgrad|dgelu" gpurun_out/gemm2cta_check.log | cut -c1-330 | tail -12
TDP_GEMM_2CTA=1 T 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_n1_2cta.json 2> gpurun_out/bench_ours_n1_2cta.err; tail -2 gpurun_out/bench_ours_n1_2cta.err; cut -c1-330 gpurun_out/bench_ours_n1_2cta.json
TR 200 29511 scripts/symm_check.py > gpurun_out/symm_check_w$N.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/symm_check_w$N.log | grep -E "ALL_OK|FAIL|Error|gemm_rs|ag_gemm" | cut -c1-420 | tail -8
TR 200 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/tp_check_w$N.log | grep -v "spin wait" | tail -6 | cut -c1-330
for impl in ours reference; do
  TR 150 29520 scripts/bench_tp.py --impl $impl > gpurun_out/bench_tp_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_tp_${impl}_w$N.log | tail -2 | cut -c1-330
done
TDP_AG_PUSH=side TR 150 29521 scripts/bench_tp.py --impl ours > gpurun_out/bench_tp_ours_side_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_tp_ours_side_w$N.log | tail -2 | cut -c1-330
