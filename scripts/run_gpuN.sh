#!/bin/bash
# multi-GPU validation + benches: bash scripts/run_gpuN.sh <N> [quick]
N=${1:-4}
mkdir -p gpurun_out
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?" >&2; }
F='Warning|warn|^$|\*\*\*|OMP_NUM'
TR 150 29511 scripts/symm_check.py > gpurun_out/symm_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/symm_check_w$N.log | grep -E "multicast|ALL_OK|FAIL|Error|gemm_rs|ag_gemm|'MiB': 256" | cut -c1-300 | tail -8
TR 150 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/tp_check_w$N.log | grep -v "spin wait" | tail -5 | cut -c1-300
TR 150 29518 scripts/ddp_debug.py > gpurun_out/ddp_debug_w$N.log 2>&1
grep -vE "$F" gpurun_out/ddp_debug_w$N.log | grep -E "^A raw|mismatching|DONE|Error" | tail -10 | cut -c1-300
TR 150 29513 scripts/engines_check.py > gpurun_out/engines_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/engines_check_w$N.log | grep -v "spin wait" | tail -8 | cut -c1-300
TR 150 29519 scripts/exposed_comm.py > gpurun_out/exposed_comm_w$N.log 2>&1; grep -E "^\{|Error" gpurun_out/exposed_comm_w$N.log | tail -1 | cut -c1-330
for impl in reference ours; do
  TR 150 29514 bench.py --impl $impl --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${impl}_n$N.json 2> gpurun_out/bench_${impl}_n$N.err; cut -c1-300 gpurun_out/bench_${impl}_n$N.json
  TR 150 29515 scripts/bench_mixed.py --impl $impl > gpurun_out/bench_mixed_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_mixed_${impl}_w$N.log | tail -2 | cut -c1-330
  TR 150 29516 scripts/bench_tp.py --impl $impl > gpurun_out/bench_tp_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_tp_${impl}_w$N.log | tail -2 | cut -c1-330
  TR 150 29517 scripts/bench_moe.py --impl $impl > gpurun_out/bench_moe_${impl}_w$N.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_moe_${impl}_w$N.log | tail -2 | cut -c1-330
done
