#!/bin/bash
# N-GPU validation of the config #3/#4 bench scripts + fused no-SP all-reduce path
N=${1:-2}
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?"; }
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?"; }
TR 200 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/tp_check_w$N.log | tail -8 | cut -c1-330
for impl in ours ours_nccl reference; do
  TR 150 29520 scripts/bench_tp.py --impl $impl > gpurun_out/bench_tp_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_tp_${impl}_w$N.log | tail -2 | cut -c1-330
done
for impl in ours ours_a2a reference; do
  TR 150 29530 scripts/bench_moe.py --impl $impl > gpurun_out/bench_moe_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_moe_${impl}_w$N.log | tail -2 | cut -c1-330
done
